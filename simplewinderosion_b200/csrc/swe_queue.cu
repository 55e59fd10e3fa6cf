// swe_queue.cu -- parallel World::erode of a single-GPU engine: the COLOUR-QUEUE scheduler.
//
// The reference flies its particles strictly one after another and relaxes the slope right after
// every step (world.hpp:47-63, wind.hpp:108-114). Flying all particles of a call in lock step with
// one deferred cascade per round (round 1's pipeline, still used by the stripe engines) exports
// +24 % more sand than the reference at 0.25 particles per cell and decorrelates the erosion field
// (0.80). tools/par_emul.c (CPU emulator on the oracle's physics) located the cause: particles
// sharing a cell within a round, the deferred cascade, and all particles having the same age at
// the same time. This scheduler removes all three:
//
//   * The cells are coloured 3 x 3 (colour = (x%3)*3 + z%3). A MICRO-ROUND serves one colour: every
//     particle standing on a cell of that colour claims its cell (one winner per cell, losers wait
//     for the colour's next turn), loads the 3x3 block into registers, flies one Wind::fly step,
//     erodes / deposits and runs World::cascade on the block at once (swe_block.cuh, bitwise the
//     reference's cascade), and stores the block back. Blocks of one colour are disjoint, so a
//     micro-round is conflict-free: plain loads and stores on the height plane, no atomics, and it
//     is equivalent to processing its particles one after another in any order -- exactly the
//     reference's order of operations (fly, erode, cascade) per particle.
//   * Particles live in 9 x 2 queues (colour x parity). After its step a particle is appended to the
//     queue of its NEW cell's colour; the queue being served is read-only during the micro-round
//     (particles that stay on the colour go to its other parity).
//   * STAGGERED INJECTION: the call's particles are not all spawned at once but spread evenly over
//     the first S micro-rounds, so that all ages are in flight together, as in the reference where a
//     cell sees young and old particles in random order.
//
// Measured with the emulator against the sequential oracle (1024^2 noise map, 0.25 particles per
// cell, tolerances of tests/test_gpu_parallel.py: exported sand +-5 %, erosion-field correlation
// > 0.95): lock-step rounds +23 % / 0.81; colour queues without stagger +8.7 % / 0.87; with S = 514
// rounds +4.5 % / 0.95, S = 1028 rounds +2.5 % / 0.97 (noise floor of two sequential runs: 0.979).
//
// Kernel: one launch per micro-round, persistent-size grid with a grid-stride loop over
// [queue entries | particles injected in this micro-round]; queue appends are aggregated per warp
// (match.any) and per block (shared counters) so the 18 global counters see one atomic per block
// and target; the last block to finish resets the served queue's counter.
#include "swe_block.cuh"
#include "swe_flight.cuh"

namespace swe {

struct QueueView {
  float4* a;            // [18][cap]  {pos.x, pos.y, pos.z, sediment}
  float4* b;            // [18][cap]  {speed.x, speed.y, speed.z, age}
  unsigned* count;      // [18] counters, 128 B apart
  unsigned* ticket;     // blocks finished in the current launch
  unsigned* dropped;    // particles lost to a full queue (must stay 0)
  int cap;
};

struct MicroRound {
  int colour;           // colour served
  unsigned pin;         // bit c = parity of colour c's input side
  unsigned index;       // micro-round number within the call (RNG counter word)
  unsigned tag;         // claim value of this micro-round (1..65535)
  int inj_first, inj_stride, inj_count, n_total;   // particles inj_first + j * inj_stride are spawned now
};

template <bool LAYERED>
__global__ void __launch_bounds__(256, 3)
colour_round_kernel(MapView m, ParConst C, block::Thr T, QueueView Q, MicroRound R, unsigned short* __restrict__ claim,
                    uint2 key, uint32_t call, const CallCtl* __restrict__ ctl, DevStats* stats) {
  pdl_sync();
  const unsigned tin = (unsigned)R.colour * 2u + ((R.pin >> R.colour) & 1u);
  const unsigned n_in = min(__ldcg(Q.count + tin * 32u), (unsigned)Q.cap);
  const unsigned total = n_in + (unsigned)R.inj_count;
  if (total == 0u) return;                        // uniform over the grid: the served counter is stable
  if (ctl) { key = make_uint2(ctl->key_x, ctl->key_y); call += ctl->call; }
  __shared__ unsigned s_cnt[18], s_base[18];
  __shared__ unsigned s_stat[4];
  __shared__ float s_lost;
  const int tid = (int)threadIdx.x;
  const unsigned lane = (unsigned)tid & 31u;
  if (tid < 4) s_stat[tid] = 0u;
  if (tid == 4) s_lost = 0.0f;
  __syncthreads();
  unsigned n_steps = 0, n_oob = 0, n_age = 0, n_new = 0;
  float lost = 0.0f;
  const int nz = m.nz;
  const float4* qa = Q.a + (size_t)tin * (size_t)Q.cap;
  const float4* qb = Q.b + (size_t)tin * (size_t)Q.cap;

  for (unsigned j0 = blockIdx.x * 256u; j0 < total; j0 += gridDim.x * 256u) {
    if (tid < 18) s_cnt[tid] = 0u;
    __syncthreads();
    const unsigned j = j0 + (unsigned)tid;
    int target = -1;
    float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < n_in) {
      ra = __ldcs(qa + j);
      rb = __ldcs(qb + j);
      const int ix = __float2int_rd(ra.x + 0.5f), iz = __float2int_rd(ra.z + 0.5f);   // wind.hpp:30 (in the map, colour served)
      const size_t ci = m.idx(ix, iz);
      // one particle per cell per micro-round; the others wait for the colour's next turn
      const unsigned short seen = *reinterpret_cast<volatile unsigned short*>(claim + ci);
      const bool won = seen != (unsigned short)R.tag && atomicCAS(claim + ci, seen, (unsigned short)R.tag) == seen;
      if (!won) {
        target = (int)(tin ^ 1u);
      } else {
        float* hp = m.height + ci;
        const bool xm = ix > 0, xp = ix + 1 < m.nx, zm = iz > 0, zp = iz + 1 < nz;
        const unsigned valid = (xm && zm ? 1u : 0u) | (xm ? 2u : 0u) | (xm && zp ? 4u : 0u) | (zm ? 8u : 0u) |
                               (zp ? 16u : 0u) | (xp && zm ? 32u : 0u) | (xp ? 64u : 0u) | (xp && zp ? 128u : 0u);
        float h[8];
        const float hc = __ldcg(hp);
#pragma unroll
        for (int k = 0; k < 8; k++)
          h[k] = ((valid >> k) & 1u) ? __ldcg(hp + block::slot_dx(k) * nz + block::slot_dz(k)) : hc;
        float rock_c = 0.0f, rk[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (LAYERED) {
          const float* rp = m.rock + ci;
          rock_c = __ldcg(rp);
#pragma unroll
          for (int k = 0; k < 8; k++)
            if ((valid >> k) & 1u) rk[k] = __ldcg(rp + block::slot_dx(k) * nz + block::slot_dz(k));
        }
        PState q;
        q.px = ra.x; q.py = ra.y; q.pz = ra.z; q.sed = ra.w;
        q.sx = rb.x; q.sy = rb.y; q.sz = rb.z; q.age = __float_as_int(rb.w);
        float collision, tx, ty, tz, tsed;
        float dh = fly_owned(C, q, hc, h[1], h[6], h[3], h[4], xm, xp, zm, zp,
                             make_uint4((uint32_t)ci, call, R.index, 0x9u), key, collision, tx, ty, tz, tsed);
        red_add_v4(m.track + ci, tsed, tx, ty, tz);                               // wind.hpp:93-96
        float c1 = hc;
        if (LAYERED && dh > 0.0f) {
          // layered extension: only the sand above the bedrock is picked up; the unmet demand grinds the
          // rock at `abrasion` efficiency (rock -> suspended sediment). The cell is owned: plain updates.
          const float take = fminf(dh, fmaxf(hc - rock_c, 0.0f));
          const float ground = C.abrasion * (dh - take);
          if (ground > 0.0f) { rock_c -= ground; m.rock[ci] = rock_c; }
          dh = take + ground;
        }
        c1 = hc - dh;                                                             // :110-112
        q.sed += dh;
        // wind.hpp:114: World::cascade at the cell the particle stood on, on the register block
        float v[8];
        int s[8];
        const unsigned moved = LAYERED ? block::cascade_sorted<true>(c1, h, valid, T, v, s, rock_c, rk)
                                       : block::cascade_sorted<false>(c1, h, valid, T, v, s, rock_c, rk);
        if (moved != 0u || dh != 0.0f) *hp = c1;
#pragma unroll
        for (int i = 0; i < 8; i++)
          if ((moved >> i) & 1u) hp[block::slot_dx(s[i]) * nz + block::slot_dz(s[i])] = v[i];
        n_steps = n_steps + 1u;
        ra = make_float4(q.px, q.py, q.pz, q.sed);
        rb = make_float4(q.sx, q.sy, q.sz, __int_as_float(q.age));
        const int jx = __float2int_rd(q.px + 0.5f), jz = __float2int_rd(q.pz + 0.5f);
        if (m.oob(jx, jz)) { n_oob++; lost += q.sed; }                            // :32-38 (the next fly() call)
        else if (q.age > C.max_age) { n_age++; lost += q.sed; }                   // :45-46
        else {
          const int cn = cell_colour(jx, jz);
          target = cn * 2 + (int)(((R.pin >> cn) & 1u) ^ (cn == R.colour ? 1u : 0u));
        }
      }
    } else if (j < total) {
      // Wind ctor + spawn (world.hpp:52-54), staggered: this micro-round's share of the call
      const int id = R.inj_first + (int)(j - n_in) * R.inj_stride;
      float px, pz;
      spawn_position_iid(id, m.nx, m.nz, key, call, px, pz);
      ra = make_float4(px, 0.0f, pz, 0.0f);
      rb = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(0));
      n_new++;
      const int jx = __float2int_rd(px + 0.5f), jz = __float2int_rd(pz + 0.5f);
      if (m.oob(jx, jz)) n_oob++;                                                 // dies in its first fly() call
      else {
        const int cn = cell_colour(jx, jz);
        target = cn * 2 + (int)(((R.pin >> cn) & 1u) ^ (cn == R.colour ? 1u : 0u));
      }
    }
    // ---- append to the target queues: warp-aggregated, then one global atomic per block and target ----
    const unsigned has = __ballot_sync(0xffffffffu, target >= 0);
    unsigned rank = 0u;
    if (target >= 0) {
      const unsigned grp = __match_any_sync(has, target);
      const int leader = __ffs(grp) - 1;
      unsigned wb = 0u;
      if ((int)lane == leader) wb = atomicAdd(&s_cnt[target], (unsigned)__popc(grp));
      wb = __shfl_sync(grp, wb, leader);
      rank = wb + (unsigned)__popc(grp & ((1u << lane) - 1u));
    }
    __syncthreads();
    if (tid < 18 && s_cnt[tid] != 0u) s_base[tid] = atomicAdd(Q.count + tid * 32, s_cnt[tid]);
    __syncthreads();
    if (target >= 0) {
      const unsigned slot = s_base[target] + rank;
      if (slot < (unsigned)Q.cap) {
        const size_t o = (size_t)target * (size_t)Q.cap + slot;
        Q.a[o] = ra;
        Q.b[o] = rb;
      } else {                                                                    // cannot happen with cap >= cycles
        atomicAdd(Q.dropped, 1u);
        lost += ra.w; n_age++;
      }
    }
  }
  // ---- counters: one set of atomics per block -----------------------------------------------------------
  for (int o = 16; o > 0; o >>= 1) {
    n_steps += __shfl_xor_sync(0xffffffffu, n_steps, o);
    n_oob += __shfl_xor_sync(0xffffffffu, n_oob, o);
    n_age += __shfl_xor_sync(0xffffffffu, n_age, o);
    n_new += __shfl_xor_sync(0xffffffffu, n_new, o);
    lost += __shfl_xor_sync(0xffffffffu, lost, o);
  }
  if (lane == 0u) {
    if (n_steps) atomicAdd(&s_stat[0], n_steps);
    if (n_oob) atomicAdd(&s_stat[1], n_oob);
    if (n_age) atomicAdd(&s_stat[2], n_age);
    if (n_new) atomicAdd(&s_stat[3], n_new);
    if (lost != 0.0f) atomicAdd(&s_lost, lost);
  }
  __syncthreads();
  if (tid == 0) {
    if (s_stat[0]) atomicAdd(&stats->steps, (unsigned long long)s_stat[0]);
    if (s_stat[1]) atomicAdd(&stats->exits_oob, (unsigned long long)s_stat[1]);
    if (s_stat[2]) atomicAdd(&stats->exits_age, (unsigned long long)s_stat[2]);
    if (s_stat[3]) atomicAdd(&stats->particles, (unsigned long long)s_stat[3]);
    if (s_lost != 0.0f) atomicAdd(&stats->sediment_lost, (double)s_lost);
    // the last block to finish empties the served queue (every block has read its length by now)
    __threadfence();
    if (atomicAdd(Q.ticket, 1u) == gridDim.x - 1u) {
      Q.count[tin * 32u] = 0u;
      *Q.ticket = 0u;
    }
  }
}


// =====================================================================================================
// Persistent form: ONE cooperative launch runs every micro-round of a World::erode call. All blocks are
// co-resident and meet at a grid barrier after each micro-round (arrival counter + a 64-bit release word
// that also carries the decision "all particles are dead" and the length of the next micro-round's
// queue, so that a block learns everything it needs from the word it polls anyway). A micro-round then
// costs one barrier (~1.5 us) instead of one dependent kernel launch (~6 us in a CUDA graph), the call ends
// exactly when the last particle dies (no empty launches, no host round trip), and the claim is one
// atomicMax on a 32-bit plane whose tags grow monotonically across calls (no per-call clearing).
// Data another block wrote in an earlier micro-round is read with ld.global.cg (L2): L1 is not coherent.
// =====================================================================================================
struct PersistCtl {
  // line 0: polled by every block
  unsigned long long release;   // [63:34] barrier generation, [33] error, [32] done, [31:0] length of the next micro-round's queue
  unsigned pad0[30];
  // line 1: arrival word, one atomic per block and micro-round: low 32 bits = blocks arrived, high 32 bits = change of
  // the number of particles in the queues (signed) -- the last block learns both from the value the atomic returns
  unsigned long long arrive;
  unsigned pad1[30];
  // line 2: written by the last block only
  int live;                     // particles in the queues
  unsigned tag_base;            // claim tags used so far (monotonic across calls)
  unsigned micro;               // micro-rounds of the last call (output)
  unsigned error;               // 1 = barrier watchdog fired, 2 = micro-round cap reached with particles alive
  unsigned long long xseq;      // stripes: micro-rounds run so far = sequence number of the cross-GPU barrier (same on every rank)
};
static_assert(sizeof(PersistCtl) == 280, "PersistCtl layout (swe_api.cu reads it back)");

__device__ __forceinline__ unsigned long long atom_add_acqrel_u64(unsigned long long* p, unsigned long long v) {
  unsigned long long old;
  asm volatile("atom.acq_rel.gpu.global.add.u64 %0, [%1], %2;" : "=l"(old) : "l"(p), "l"(v) : "memory");
  return old;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

#ifndef SWE_RELEASE_COPIES
#define SWE_RELEASE_COPIES 1u    // (16 copies on separate L2 slices were measured: no change)
#endif
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// copy 0 is PersistCtl::release itself; copies 1.. live 128 bytes apart behind the stripe mailbox
__device__ __forceinline__ unsigned long long* release_copy(PersistCtl* ctl, unsigned g) {
  return g == 0u ? &ctl->release
                 : reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(ctl) + SWE_GROUP_OFFSET) + g * 16u;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// L2 residency steering (createpolicy + .L2::cache_hint): the height plane and the claim lattice are what
// every step re-reads (evict_last); particle records are read exactly once and the track plane is only
// ever added to (evict_first), so they should not push the heights out of the 126 MB L2.
#ifndef SWE_L2_HINTS
#define SWE_L2_HINTS 1
#endif
__device__ __forceinline__ unsigned long long l2_policy_keep() {
  unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ unsigned long long l2_policy_stream() {
  unsigned long long p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ float ld_l2(const float* a, unsigned long long pol) {
#if SWE_L2_HINTS
  float v; asm volatile("ld.global.cg.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(a), "l"(pol) : "memory"); return v;
#else
  return __ldcg(a);
#endif
}
__device__ __forceinline__ float4 ld_l2(const float4* a, unsigned long long pol) {
#if SWE_L2_HINTS
  float4 v;
  asm volatile("ld.global.cg.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a), "l"(pol) : "memory");
  return v;
#else
  return __ldcg(a);
#endif
}
__device__ __forceinline__ void st_l2(float* a, float v, unsigned long long pol) {
#if SWE_L2_HINTS
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(a), "f"(v), "l"(pol) : "memory");
#else
  *a = v;
#endif
}
__device__ __forceinline__ void red_add_v4_l2(float4* a, float x, float y, float z, float w, unsigned long long pol) {
#if SWE_L2_HINTS
  asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(a), "f"(x), "f"(y), "f"(z), "f"(w), "l"(pol) : "memory");
#else
  red_add_v4(a, x, y, z, w);
#endif
}
__device__ __forceinline__ unsigned atom_max_l2(unsigned* a, unsigned v, unsigned long long pol) {
#if SWE_L2_HINTS
  unsigned o; asm volatile("atom.global.max.L2::cache_hint.u32 %0, [%1], %2, %3;" : "=r"(o) : "l"(a), "r"(v), "l"(pol) : "memory"); return o;
#else
  return atomicMax(a, v);
#endif
}

// ---- multi-GPU row stripes ---------------------------------------------------------------------------------
// Every rank runs the same micro-round sequence; the grid barrier of a stripe kernel extends over all ranks
// (mailbox words written into every rank's memory over NVLink, system-scope release/acquire), so the colouring
// holds GLOBALLY: a particle in a stripe's first or last row reads and writes the neighbour's rows of its 3x3
// block directly in the neighbour's memory (peer loads / stores) with the same exclusive ownership as inside
// one GPU -- no ghost rows, no halo exchange, and the result has the single-GPU semantics. A particle whose
// next cell lies in a neighbour's rows is appended to the neighbour's colour queue (system-scope atomic for
// the slot, peer stores for the record).
__device__ __forceinline__ float ld_peer(const float* a) {
  float v; asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(a) : "memory"); return v;
}
__device__ __forceinline__ void st_peer(float* a, float v) {
  asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_h(const float* a, bool remote, unsigned long long pol) { return remote ? ld_peer(a) : ld_l2(a, pol); }
__device__ __forceinline__ void st_h(float* a, float v, bool remote, unsigned long long pol) { if (remote) st_peer(a, v); else st_l2(a, v, pol); }

__device__ __forceinline__ unsigned pin_mask(unsigned k) {
  // colour c has been served floor((k - c + 8) / 9) times before micro-round k
  unsigned pin = 0u;
#pragma unroll
  for (unsigned c = 0; c < 9; c++) pin |= (((k + 8u - c) / 9u) & 1u) << c;
  return pin;
}

#ifndef SWE_PERSIST_THREADS
#define SWE_PERSIST_THREADS 256   // threads per block of the persistent kernel
#endif
#ifndef SWE_PERSIST_BLOCKS
#define SWE_PERSIST_BLOCKS (1024 / SWE_PERSIST_THREADS)   // resident blocks per SM the register budget is tuned for (<= 64 registers)
#endif
#ifndef SWE_WARP_APPEND
#define SWE_WARP_APPEND 0         // 1 = queue appends reserved per warp instead of per block (measured on B200: 1.5x SLOWER, the 18 counters become hot)
#endif
#ifndef SWE_STRIPE_SYS_ATOMICS
#define SWE_STRIPE_SYS_ATOMICS 1  // stripes: a rank reserves slots in its OWN queues with system-scope atomics too (the neighbours' hand-offs are)
#endif
#ifndef SWE_PREFETCH_NEXT
#define SWE_PREFETCH_NEXT 0       // 1 = a particle prefetches the rows of its NEXT cell into L2 when its step ends (measured: no gain)
#endif
#ifndef SWE_MULTI_STEP
#define SWE_MULTI_STEP 1u         // steps a particle may take per turn while it stays on its cell (4: 32 % fewer micro-rounds and a better
                                  // match with the oracle, but 1.3-1.6x SLOWER on B200: one repeating lane holds its warp, and every
                                  // micro-round, for a second dependent chain)
#endif
#ifndef SWE_PREFETCH
#define SWE_PREFETCH 0            // 1 = blocks waiting at the grid barrier prefetch the next micro-round's heights into L2 (measured on B200: 8 % SLOWER)
#endif
constexpr unsigned PNT = SWE_PERSIST_THREADS;
__device__ __forceinline__ void prefetch_l2_keep(const void* a) {
  asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(a));
}
__device__ __forceinline__ void prefetch_l2(const void* a) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
}
template <bool LAYERED, bool STRIPE>
__global__ void __launch_bounds__(SWE_PERSIST_THREADS, SWE_PERSIST_BLOCKS)
colour_persistent_kernel(MapView m, ParConst C, block::Thr T, QueueView Q, unsigned* __restrict__ claim, uint2 key,
                         uint32_t call, int n_total, unsigned stagger, unsigned max_micro, PersistCtl* ctl,
                         DevStats* stats, const StripeQueues S) {
#if !SWE_WARP_APPEND
  __shared__ unsigned s_cnt[18], s_base[18];
#endif
  __shared__ unsigned s_stat[4];
  __shared__ float s_lost;
  __shared__ unsigned long long s_rel;
  __shared__ int s_delta;
  const int tid = (int)threadIdx.x;
  const unsigned lane = (unsigned)tid & 31u;
  if (tid < 4) s_stat[tid] = 0u;
  if (tid == 4) { s_lost = 0.0f; s_delta = 0; }
  if (tid == 0) s_rel = ld_relaxed_u64(&ctl->release);
  __syncthreads();
  const unsigned tag0 = *reinterpret_cast<volatile unsigned*>(&ctl->tag_base);   // stable until the final barrier
  const unsigned long long xseq0 = STRIPE ? *reinterpret_cast<volatile unsigned long long*>(&ctl->xseq) : 0ull;
  const int lat0 = m.x0 / 3;                               // first lattice row of the local claim table
  unsigned long long rel = s_rel;
  unsigned n_steps = 0, n_oob = 0, n_age = 0, n_new = 0;
  float lost = 0.0f;
  const int nz = m.nz;
  const int claim_pitch = (nz + 2) / 3;
  const unsigned long long keep = l2_policy_keep(), stream = l2_policy_stream();
  unsigned n_in = 0u;                                     // micro-round 0: all queues are empty
  // per-call uniforms kept out of the micro-round loop (they showed up as ~130 instructions per particle step)
  const long long seam_rows = (long long)(m.nx - 1) * (long long)nz;     // periodic: offset from the first to the last row
  const unsigned inj_q = stagger ? (unsigned)n_total / stagger : 0u, inj_r = stagger ? (unsigned)n_total % stagger : 0u;
  unsigned pin = 0u, colour = 0u;                          // parity bits of the queues' input sides; colour served = k % 9

  for (unsigned k = 0;; k++) {
    const unsigned tin = colour * 2u + ((pin >> colour) & 1u);
    const unsigned tag = tag0 + k + 1u;
    const unsigned inj_count = k < stagger ? inj_q + (k < inj_r ? 1u : 0u) : 0u;   // ids k + j * stagger < n_total
    const unsigned total = n_in + inj_count;
    const float4* qa = Q.a + (size_t)tin * (size_t)Q.cap;
    const float4* qb = Q.b + (size_t)tin * (size_t)Q.cap;
    int delta = 0;                                        // particles this thread added to / removed from the queues

    for (unsigned j0 = blockIdx.x * PNT; j0 < total; j0 += gridDim.x * PNT) {
#if !SWE_WARP_APPEND
      if (tid < 18) s_cnt[tid] = 0u;
      __syncthreads();
#endif
      const unsigned j = j0 + (unsigned)tid;
      int target = -1;
      float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < n_in) {
        ra = ld_l2(qa + j, stream);
        rb = ld_l2(qb + j, stream);
        int ix = __float2int_rd(ra.x + 0.5f), iz = __float2int_rd(ra.z + 0.5f);         // wind.hpp:30
        if (m.periodic) { ix = MapView::wrap(ix, m.nx); iz = MapView::wrap(iz, nz); }     // cellpool.h:102-110 (positions keep growing)
        const size_t ci = m.idx(ix, iz);
        float* hp = m.height + ci;
        const bool per = m.periodic != 0;
        const bool xm = per || ix > 0, xp = per || ix + 1 < m.nx, zm = per || iz > 0, zp = per || iz + 1 < nz;
        const unsigned valid = (xm && zm ? 1u : 0u) | (xm ? 2u : 0u) | (xm && zp ? 4u : 0u) | (zm ? 8u : 0u) |
                               (zp ? 16u : 0u) | (xp && zm ? 32u : 0u) | (xp ? 64u : 0u) | (xp && zp ? 128u : 0u);
        // rows / columns of the 3x3 block: a stripe reads / writes a neighbour's row in the neighbour's memory, a
        // periodic map wraps at the seam
        const bool rem_m = STRIPE && xm && ix - 1 < S.own_lo, rem_p = STRIPE && xp && ix + 1 >= S.own_hi;
        float* row_m = rem_m ? S.below_h + (size_t)(ix - 1 - S.below_x0) * (size_t)nz + iz
                             : ((per && ix == 0) ? hp + seam_rows : hp - nz);
        float* row_p = rem_p ? S.above_h + (size_t)(ix + 1 - S.above_x0) * (size_t)nz + iz
                             : ((per && ix == m.nx - 1) ? hp - seam_rows : hp + nz);
        const int zo_m = (per && iz == 0) ? nz - 1 : -1, zo_p = (per && iz == nz - 1) ? 1 - nz : 1;
        // one claim word per 3x3 lattice cell: the cells of the served colour map one-to-one onto it, the tags
        // grow with every micro-round, so the table (1/9 of a plane: L2-resident) serves all colours.
        // The block is loaded while the claim is in flight (harmless if the claim is lost).
        const unsigned seen = atom_max_l2(claim + (size_t)(ix / 3 - lat0) * (size_t)claim_pitch + (size_t)(iz / 3), tag, keep);
        float h[8];
        float hc = ld_l2(hp, keep);
#pragma unroll
        for (int q8 = 0; q8 < 8; q8++) {
          const int dx = block::slot_dx(q8), dz = block::slot_dz(q8);
          const float* src = (dx < 0 ? row_m : (dx > 0 ? row_p : hp)) + (dz < 0 ? zo_m : (dz > 0 ? zo_p : 0));
          h[q8] = ((valid >> q8) & 1u) ? ld_h(src, dx < 0 ? rem_m : (dx > 0 ? rem_p : false), keep) : hc;
        }
        if (seen >= tag) {
          target = (int)(tin ^ 1u);                        // the cell is taken: wait for the colour's next turn
        } else {
          float rock_c = 0.0f, rk[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (LAYERED) {
            const float* rp = m.rock + ci;
            rock_c = ld_l2(rp, keep);
#pragma unroll
            for (int q8 = 0; q8 < 8; q8++)
              if ((valid >> q8) & 1u) rk[q8] = ld_l2(rp + block::slot_dx(q8) * nz + block::slot_dz(q8), keep);
          }
          PState q;
          q.px = ra.x; q.py = ra.y; q.pz = ra.z; q.sed = ra.w;
          q.sx = rb.x; q.sy = rb.y; q.sz = rb.z; q.age = __float_as_int(rb.w);
          // A particle that is still on its cell after a step keeps the claim and steps again at once (it owns the
          // cell and the block for the whole micro-round; in the reference nothing comes between two steps of a
          // particle either): up to SWE_MULTI_STEP steps per turn instead of a nine-micro-round wait per step for the
          // slow early part of a flight. Emulated (tools/par_emul.c): 22 % fewer micro-rounds at 1024^2 / 0.25 per
          // cell and a better match with the sequential oracle (exported sand +0.9 % instead of +2.5 %).
#pragma unroll 1
          for (unsigned rep = 0;; rep++) {
            if (rep > 0u) {                                // the block as this thread left it
              hc = ld_l2(hp, keep);
#pragma unroll
              for (int q8 = 0; q8 < 8; q8++) {
                const int dx = block::slot_dx(q8), dz = block::slot_dz(q8);
                const float* src = (dx < 0 ? row_m : (dx > 0 ? row_p : hp)) + (dz < 0 ? zo_m : (dz > 0 ? zo_p : 0));
                h[q8] = ((valid >> q8) & 1u) ? ld_h(src, dx < 0 ? rem_m : (dx > 0 ? rem_p : false), keep) : hc;
              }
            }
            float collision, tx, ty, tz, tsed;
            float dh = fly_owned(C, q, hc, h[1], h[6], h[3], h[4], xm, xp, zm, zp,
                                 make_uint4((uint32_t)ci, call, k, 0x9u + 16u * rep), key, collision, tx, ty, tz, tsed);
            red_add_v4_l2(m.track + ci, tsed, tx, ty, tz, stream);                  // wind.hpp:93-96
            if (LAYERED && dh > 0.0f) {
              const float take = fminf(dh, fmaxf(hc - rock_c, 0.0f));
              const float ground = C.abrasion * (dh - take);
              if (ground > 0.0f) { rock_c -= ground; m.rock[ci] = rock_c; }
              dh = take + ground;
            }
            float c1 = hc - dh;                                                     // :110-112
            q.sed += dh;
            float v[8];
            int s[8];
            const unsigned moved = LAYERED ? block::cascade_sorted<true>(c1, h, valid, T, v, s, rock_c, rk)   // :114
                                           : block::cascade_sorted<false>(c1, h, valid, T, v, s, rock_c, rk);
            if (moved != 0u || dh != 0.0f) st_l2(hp, c1, keep);
#pragma unroll
            for (int i = 0; i < 8; i++)
              if ((moved >> i) & 1u) {
                const int dx = block::slot_dx(s[i]), dz = block::slot_dz(s[i]);
                float* dst = (dx < 0 ? row_m : (dx > 0 ? row_p : hp)) + (dz < 0 ? zo_m : (dz > 0 ? zo_p : 0));
                st_h(dst, v[i], dx < 0 ? rem_m : (dx > 0 ? rem_p : false), keep);
              }
            n_steps++;
            int jx = __float2int_rd(q.px + 0.5f), jz = __float2int_rd(q.pz + 0.5f);
            if (per) { jx = MapView::wrap(jx, m.nx); jz = MapView::wrap(jz, nz); }
            if (m.oob(jx, jz)) { n_oob++; lost += q.sed; delta--; break; }          // :32-38 (the next fly() call)
            if (q.age > C.max_age) { n_age++; lost += q.sed; delta--; break; }      // :45-46
            if (jx == ix && jz == iz && rep + 1u < SWE_MULTI_STEP) continue;        // still on its cell: next step now
            ra = make_float4(q.px, q.py, q.pz, q.sed);
            rb = make_float4(q.sx, q.sy, q.sz, __int_as_float(q.age));
            const unsigned cn = (unsigned)cell_colour(jx, jz);
            target = (int)(cn * 2u + (((pin >> cn) & 1u) ^ (cn == colour ? 1u : 0u)));
#if SWE_PREFETCH_NEXT
            // the particle knows where it will stand at its next turn (3-9 micro-rounds from now): pull that cell's
            // rows into L2 now, off the critical path of that micro-round (no load, no register: prefetch only)
            if (!STRIPE || (jx > S.own_lo && jx + 1 < S.own_hi)) {
              const float* nh = m.height + m.idx(jx, jz);
              prefetch_l2_keep(nh);
              if (jx > 0) prefetch_l2_keep(nh - nz);
              if (jx + 1 < m.nx) prefetch_l2_keep(nh + nz);
            }
#endif
            if (STRIPE && (jx < S.own_lo || jx >= S.own_hi)) {
              // hand the particle to the owner of its next cell: append to the neighbour's queue of that colour
              const bool dn = jx < S.own_lo;
              unsigned* pc = (dn ? S.below_count : S.above_count) + target * 32;
              const int pcap = dn ? S.below_cap : S.above_cap;
              const unsigned slot = atomicAdd_system(pc, 1u);
              if (slot < (unsigned)pcap) {
                const size_t o = (size_t)target * (size_t)pcap + slot;
                (dn ? S.below_a : S.above_a)[o] = ra;
                (dn ? S.below_b : S.above_b)[o] = rb;
              } else {
                atomicAdd(Q.dropped, 1u);
                lost += ra.w; n_age++; delta--;
              }
              target = -1;                                 // (the global live count is the sum over the ranks)
            }
            break;
          }
        }
      } else if (j < total) {
        const int id = (int)k + (int)(j - n_in) * (int)stagger;                   // world.hpp:52-54, staggered
        float px, pz;
        spawn_position_iid(id, m.nx, m.nz, key, call, px, pz);
        ra = make_float4(px, 0.0f, pz, 0.0f);
        rb = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(0));
        const int jx = __float2int_rd(px + 0.5f), jz = __float2int_rd(pz + 0.5f);
        // a stripe materialises the particles whose first cell lies in its rows (every rank evaluates the whole
        // injection list of the micro-round; a cell beyond the last row belongs to the last rank)
        const bool mine = !STRIPE || (jx >= S.own_lo && (jx < S.own_hi || S.own_hi == m.nx));
        if (mine) {
          n_new++;
          if (m.oob(jx, jz)) n_oob++;
          else {
            delta++;
            const unsigned cn = (unsigned)cell_colour(jx, jz);
            target = (int)(cn * 2u + (((pin >> cn) & 1u) ^ (cn == colour ? 1u : 0u)));
          }
        }
      }
      const unsigned has = __ballot_sync(0xffffffffu, target >= 0);
#if SWE_WARP_APPEND
      if (target >= 0) {
        // one reservation per warp and target queue: the lanes going to the same queue elect a leader
        const unsigned grp = __match_any_sync(has, target);
        const int leader = __ffs(grp) - 1;
        unsigned wb = 0u;
        if ((int)lane == leader)
          wb = STRIPE ? atomicAdd_system(Q.count + target * 32, (unsigned)__popc(grp)) : atomicAdd(Q.count + target * 32, (unsigned)__popc(grp));
        wb = __shfl_sync(grp, wb, leader);
        const unsigned slot = wb + (unsigned)__popc(grp & ((1u << lane) - 1u));
#else
      unsigned rank = 0u;
      if (target >= 0) {
        const unsigned grp = __match_any_sync(has, target);
        const int leader = __ffs(grp) - 1;
        unsigned wb = 0u;
        if ((int)lane == leader) wb = atomicAdd(&s_cnt[target], (unsigned)__popc(grp));
        wb = __shfl_sync(grp, wb, leader);
        rank = wb + (unsigned)__popc(grp & ((1u << lane) - 1u));
      }
      __syncthreads();
      if (tid < 18 && s_cnt[tid] != 0u)
        s_base[tid] = (STRIPE && SWE_STRIPE_SYS_ATOMICS) ? atomicAdd_system(Q.count + tid * 32, s_cnt[tid]) : atomicAdd(Q.count + tid * 32, s_cnt[tid]);
      __syncthreads();
      if (target >= 0) {
        const unsigned slot = s_base[target] + rank;
#endif
        if (slot < (unsigned)Q.cap) {
          const size_t o = (size_t)target * (size_t)Q.cap + slot;
          Q.a[o] = ra;
          Q.b[o] = rb;
        } else {
          atomicAdd(Q.dropped, 1u);
          lost += ra.w; n_age++; delta--;
        }
      }
    }

    // ---- grid barrier ---------------------------------------------------------------------------------------
    if (delta != 0) atomicAdd(&s_delta, delta);
    __syncthreads();
    if (tid == 0) {
      const int d = s_delta;
      s_delta = 0;
      // release: this block's appends and height stores (ordered before by the block barrier) become visible
      // to whoever acquires the arrival word; acquire: the last block sees everybody else's
      const unsigned long long old = atom_add_acqrel_u64(&ctl->arrive, 1ull + ((unsigned long long)(long long)d << 32));
      if ((unsigned)(old & 0xffffffffull) == gridDim.x - 1u) {
        int live = ctl->live + (int)(unsigned)((old + ((unsigned long long)(long long)d << 32)) >> 32);
        ctl->live = live;
        ctl->arrive = 0ull;
        Q.count[tin * 32u] = 0u;                           // the served queue is empty again
        const unsigned kn = k + 1u;
        bool timed_out = false;
        if (STRIPE) {
          // Cross-GPU barrier. Every rank publishes (sequence number, local live count) of the micro-round it has
          // finished into every rank's mailbox ring (16 micro-rounds deep, 8 ranks per slot: one 64-byte line) and
          // then waits for its two NEIGHBOURS only -- they are the only ranks whose boundary rows and queue
          // appends it depends on, so a slow rank delays the others one hop per micro-round instead of at once.
          // Ranks can therefore drift up to nranks - 1 micro-rounds apart; the decision "every particle is dead"
          // is taken from the words of micro-round kn - nranks, which have all arrived by then (no waiting), so
          // every rank leaves the loop at the same micro-round.
          const unsigned long long seq = xseq0 + (unsigned long long)kn;
          const unsigned long long word = (seq << 32) | (unsigned long long)(unsigned)live;
          const int slot = (int)(seq & 15ull);
          // ONE system-scope fence orders everything this GPU wrote in the micro-round (observed through the
          // arrival word) before the mailbox words; the words themselves are then plain system-scope stores
#ifndef SWE_EXPERIMENT_NO_SYS_FENCE   // (measurement only: how much of the cross-GPU barrier is the fence)
          asm volatile("fence.acq_rel.sys;" ::: "memory");
#endif
          for (int i = 0; i < S.nranks; i++) st_relaxed_sys_u64(S.mailbox[i] + slot * 8 + S.rank, word);
          const unsigned long long* mine = S.mailbox[S.rank];
          const long long t0 = clock64();
          for (int nb = S.rank - 1; nb <= S.rank + 1 && !timed_out; nb += 2) {
            if (nb < 0 || nb >= S.nranks) continue;
            unsigned spins = 0;
            while ((ld_acquire_sys_u64(mine + slot * 8 + nb) >> 32) != (seq & 0xffffffffull))
              if ((++spins & 1023u) == 0u && clock64() - t0 > 16000000000LL) { timed_out = true; break; }   // ~8 s
          }
          live = 1;                                          // not decided yet
          if (kn > (unsigned)S.nranks && kn - (unsigned)S.nranks >= stagger) {     // (nothing to decide while particles are still injected)
            const unsigned long long seq_t = seq - (unsigned long long)S.nranks;
            const int slot_t = (int)(seq_t & 15ull);
            int total = 0;
            // the eight words of a ring slot share one 64-byte line: four independent 16-byte loads, one latency
            unsigned long long w[8];
#pragma unroll
            for (int i = 0; i < 8; i += 2)
              asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w[i]), "=l"(w[i + 1]) : "l"(mine + slot_t * 8 + i) : "memory");
#pragma unroll
            for (int i = 0; i < 8; i++) {
              if (i >= S.nranks) break;
              unsigned spins = 0;
              while ((w[i] >> 32) != (seq_t & 0xffffffffull)) {                     // (cannot happen: ranks drift < nranks)
                w[i] = ld_acquire_sys_u64(mine + slot_t * 8 + i);
                if ((++spins & 1023u) == 0u && clock64() - t0 > 16000000000LL) { timed_out = true; break; }
              }
              total += (int)(unsigned)(w[i] & 0xffffffffull);
            }
            // all particles were dead nranks micro-rounds ago (and all had been injected): nothing has moved since
            live = total == 0 ? 0 : 1;
          }
        }
        const unsigned cn = colour == 8u ? 0u : colour + 1u;
        const unsigned tn = cn * 2u + (((pin ^ (1u << colour)) >> cn) & 1u);
        const unsigned next_n = tn == tin ? 0u : min(__ldcg(Q.count + tn * 32u), (unsigned)Q.cap);
        unsigned done = (live == 0 && (STRIPE || kn >= stagger)) ? 1u : 0u;
        if (!done && kn >= max_micro) { done = 1u; ctl->error = 2u; }
        if (timed_out) { done = 1u; ctl->error = 1u; }
        if (done) { ctl->micro = kn; ctl->tag_base = tag0 + kn + 2u; if (STRIPE) ctl->xseq = xseq0 + (unsigned long long)kn; }
        const unsigned long long gen = (rel >> 34) + 1ull;
        rel = (gen << 34) | ((unsigned long long)done << 32) | (unsigned long long)next_n;
        // the release word is fanned out to 16 copies on separate 128-byte lines (separate L2 slices): hundreds of
        // blocks polling ONE address saturate its slice and delay the very store they are waiting for
        __threadfence();
        for (unsigned g = 1; g < SWE_RELEASE_COPIES; g++) st_relaxed_u64(release_copy(ctl, g), rel);
        st_release_u64(&ctl->release, rel);
      } else {
        const unsigned long long gen = rel >> 34;
        const long long t0 = clock64();
        unsigned long long now;
        unsigned spins = 0;
        const unsigned long long* my_rel = release_copy(ctl, blockIdx.x % SWE_RELEASE_COPIES);
        while (((now = ld_acquire_u64(my_rel)) >> 34) == gen) {
          __nanosleep(40);
          if ((++spins & 1023u) == 0u && clock64() - t0 > 8000000000LL) {   // ~4 s watchdog: never hang the device
            ctl->error = 1u; now = ((gen + 1ull) << 34) | (3ull << 32); break;
          }
        }
        rel = now;
      }
      s_rel = rel;
    }
#if SWE_PREFETCH
    else if (tid >= 32) {
      // a block that waits at the grid barrier pulls the heights its threads will need in the NEXT micro-round into
      // L2 (the entries already in that queue; addresses only -- nothing is consumed, so a record that is still
      // being written by a slower block is harmless)
      const unsigned cn = colour == 8u ? 0u : colour + 1u;
      const unsigned tn = cn * 2u + (((pin ^ (1u << colour)) >> cn) & 1u);
      const unsigned jn = blockIdx.x * PNT + (unsigned)tid;
      if (jn < min(__ldcg(Q.count + tn * 32u), (unsigned)Q.cap)) {
        const size_t o = (size_t)tn * (size_t)Q.cap + jn;
        const float4 pa = __ldcg(Q.a + o);
        prefetch_l2(Q.b + o);
        const int ix = __float2int_rd(pa.x + 0.5f), iz = __float2int_rd(pa.z + 0.5f);
        const int x_lo = STRIPE ? S.own_lo : 0, x_hi = STRIPE ? S.own_hi : m.nx;
        if (ix >= x_lo && ix < x_hi && iz >= 0 && iz < nz) {
          const float* hp = m.height + m.idx(ix, iz);
          prefetch_l2_keep(hp);
          if (ix - 1 >= x_lo) prefetch_l2_keep(hp - nz);
          if (ix + 1 < x_hi) prefetch_l2_keep(hp + nz);
          prefetch_l2(m.track + m.idx(ix, iz));
        }
      }
    }
#endif
    __syncthreads();
    rel = s_rel;
    n_in = (unsigned)(rel & 0xffffffffull);
    if ((rel >> 32) & 1ull) break;
    pin ^= 1u << colour;                                   // the served colour's input side flips
    colour = colour == 8u ? 0u : colour + 1u;
  }

  for (int o = 16; o > 0; o >>= 1) {
    n_steps += __shfl_xor_sync(0xffffffffu, n_steps, o);
    n_oob += __shfl_xor_sync(0xffffffffu, n_oob, o);
    n_age += __shfl_xor_sync(0xffffffffu, n_age, o);
    n_new += __shfl_xor_sync(0xffffffffu, n_new, o);
    lost += __shfl_xor_sync(0xffffffffu, lost, o);
  }
  if (lane == 0u) {
    if (n_steps) atomicAdd(&s_stat[0], n_steps);
    if (n_oob) atomicAdd(&s_stat[1], n_oob);
    if (n_age) atomicAdd(&s_stat[2], n_age);
    if (n_new) atomicAdd(&s_stat[3], n_new);
    if (lost != 0.0f) atomicAdd(&s_lost, lost);
  }
  __syncthreads();
  if (tid == 0) {
    if (s_stat[0]) atomicAdd(&stats->steps, (unsigned long long)s_stat[0]);
    if (s_stat[1]) atomicAdd(&stats->exits_oob, (unsigned long long)s_stat[1]);
    if (s_stat[2]) atomicAdd(&stats->exits_age, (unsigned long long)s_stat[2]);
    if (s_stat[3]) atomicAdd(&stats->particles, (unsigned long long)s_stat[3]);
    if (s_lost != 0.0f) atomicAdd(&stats->sediment_lost, (double)s_lost);
  }
}

// resident blocks per SM of the persistent kernel (0 = the device cannot run it)
int colour_persistent_blocks_per_sm(bool layered) {
  int n = 0;
  const cudaError_t err = layered
      ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, colour_persistent_kernel<true, false>, SWE_PERSIST_THREADS, 0)
      : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, colour_persistent_kernel<false, false>, SWE_PERSIST_THREADS, 0);
  return err == cudaSuccess ? n : 0;
}

int colour_persistent_blocks_per_sm_stripe() {
  int n = 0;
  const cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, colour_persistent_kernel<false, true>, SWE_PERSIST_THREADS, 0);
  return err == cudaSuccess ? n : 0;
}

cudaError_t launch_colour_persistent(const MapView& m, const swe_params& p, const ColourQueues& cq, int n_total,
                                     unsigned stagger, unsigned max_micro, uint64_t seed, uint32_t call,
                                     DevStats* stats, cudaStream_t st, int grid_blocks, const StripeQueues* stripe,
                                     bool cooperative) {
  uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  block::Thr T;
  T.ortho = sqrtf(1.0f) * p.maxdiff * 1.0f;
  T.diag = sqrtf(2.0f) * p.maxdiff * 1.0f;
  T.settling = p.settling;
  QueueView Q;
  Q.a = cq.a; Q.b = cq.b; Q.count = cq.count; Q.ticket = cq.count + 18 * 32; Q.dropped = cq.count + 19 * 32;
  Q.cap = cq.cap;
  MapView mv = m;
  ParConst C = make_const(p);
  unsigned* claim = cq.claim32;
  PersistCtl* ctl = reinterpret_cast<PersistCtl*>(cq.persist);
  StripeQueues sq{};
  if (stripe) sq = *stripe;
  void* args[] = {&mv, &C, &T, &Q, &claim, &key, &call, &n_total, &stagger, &max_micro, &ctl, &stats, &sq};
  const void* fn = stripe ? (m.rock ? reinterpret_cast<const void*>(&colour_persistent_kernel<true, true>)
                                    : reinterpret_cast<const void*>(&colour_persistent_kernel<false, true>))
                 : m.rock ? reinterpret_cast<const void*>(&colour_persistent_kernel<true, false>)
                          : reinterpret_cast<const void*>(&colour_persistent_kernel<false, false>);
  // several logical stripes on ONE GPU (tests): their kernels must run side by side, which plain launches of
  // grids sized to a share of the device do; a cooperative launch is what guarantees co-residency otherwise
  if (!cooperative) return cudaLaunchKernel(fn, dim3(grid_blocks), dim3(SWE_PERSIST_THREADS), args, 0, st);
  return cudaLaunchCooperativeKernel(fn, dim3(grid_blocks), dim3(SWE_PERSIST_THREADS), args, 0, st);
}

void launch_colour_round(const MapView& m, const swe_params& p, const ColourQueues& cq, const MicroRoundArgs& r,
                         uint64_t seed, uint32_t call, const CallCtl* ctl, DevStats* stats, cudaStream_t st, bool pdl,
                         int grid_blocks) {
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  block::Thr T;
  T.ortho = sqrtf(1.0f) * p.maxdiff * 1.0f;   // host IEEE fp32: the values exact::cascade_thresholds computes
  T.diag = sqrtf(2.0f) * p.maxdiff * 1.0f;
  T.settling = p.settling;
  QueueView Q;
  Q.a = cq.a; Q.b = cq.b; Q.count = cq.count; Q.ticket = cq.count + 18 * 32; Q.dropped = cq.count + 19 * 32;
  Q.cap = cq.cap;
  MicroRound R;
  R.colour = r.colour; R.pin = r.pin; R.index = r.index; R.tag = r.tag;
  R.inj_first = r.inj_first; R.inj_stride = r.inj_stride; R.inj_count = r.inj_count; R.n_total = r.n_total;
  if (m.rock)
    launch_chain(colour_round_kernel<true>, dim3(grid_blocks), dim3(256), 0, st, pdl, m, make_const(p), T, Q, R, cq.claim, key,
                 call, ctl, stats);
  else
    launch_chain(colour_round_kernel<false>, dim3(grid_blocks), dim3(256), 0, st, pdl, m, make_const(p), T, Q, R, cq.claim, key,
                 call, ctl, stats);
}

int colour_persistent_threads() { return SWE_PERSIST_THREADS; }

void preload_queue() {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, colour_round_kernel<false>);
  cudaFuncGetAttributes(&a, colour_round_kernel<true>);
  cudaFuncGetAttributes(&a, colour_persistent_kernel<false, false>);
  cudaFuncGetAttributes(&a, colour_persistent_kernel<true, false>);
  cudaFuncGetAttributes(&a, colour_persistent_kernel<false, true>);
  cudaFuncGetAttributes(&a, colour_persistent_kernel<true, true>);
}

}  // namespace swe
