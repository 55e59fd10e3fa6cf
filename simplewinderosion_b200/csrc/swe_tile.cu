// swe_tile.cu -- parallel World::erode: the TILE-RESIDENT scheduler (single-GPU engines, and the stripes of a
// multi-GPU map: TileStripe below).
//
// Same per-particle semantics as the colour-queue scheduler (swe_queue.cu): a particle claims its cell for a
// micro-round of its cell's colour, flies ONE Wind::fly step (wind.hpp:28-112) against the current heights,
// erodes / deposits, and runs World::cascade on its 3x3 block at once (wind.hpp:114) -- the reference's order of
// operations -- and the call's particles are injected staggered over S rounds. What changes is WHERE a
// micro-round is synchronised. The colour-queue kernel meets at a grid barrier after every micro-round (13,500
// dependent ~21 us steps per headline call: latency-bound, 2.5x the algorithmic DRAM traffic in scattered
// 32-byte sectors). Here the map is cut into T x T tiles and ONE CTA owns one tile for a SUPER-ROUND of K rounds
// (9 K micro-rounds behind __syncthreads): the tile's heights live in shared memory and are read from / written
// to the height plane once per super-round with coalesced rows; the tile's particle records are one contiguous
// segment that only this CTA touches (L2-resident for the super-round; measured on B200: keeping them in shared
// memory costs 35 bytes per particle and lowers the number of resident CTAs -- 4-7 % slower); independent CTAs
// hide each other's latency.
//
// Conflict-freedom without halos or atomics: a particle is served only while its cell lies in the INTERIOR of
// the tile (>= 1 cell from the tile's edges; map edges excepted), so the 3x3 block it reads and writes never
// leaves the tile and tiles never touch each other's cells. A particle that reaches the border ring -- or
// leaves the tile in one step -- is PARKED until the next super-round, whose tiling is shifted by T/3 in both
// directions. Three tilings take turns; a cell lies on the x-ring of at most one of them and on the z-ring of
// at most one, so every cell is interior at least every third super-round. T and the shifts are multiples of
// 3: the tile-local colouring is the global one and one claim word per 3x3 lattice cell serves all colours.
// Every schedule this produces is a sequentially consistent interleaving of whole particle steps, like the
// colour queues'. tools/par_emul.c (mode D) runs this schedule on the oracle's physics on the CPU: against
// the one-after-another reference it is statistically as close as the colour queues.
//
// Per super-round s (host loop, three launches):
//   tile_fly_kernel   one CTA per tile of tiling s: load tile + its particle segment, 9 K micro-rounds, store;
//                     counts, per tile of tiling s+1, the surviving particles (and, in extra CTAs, the
//                     particles injected for s+1)
//   tile_scan_kernel  exclusive prefix sum of those counts = segment bases of tiling s+1
//   tile_scatter_kernel  moves every surviving record (and the new particles) into its tile's segment
#include "swe_block.cuh"
#include "swe_flight.cuh"

namespace swe {

#ifndef SWE_TILE_ALWAYS_RNG
#define SWE_TILE_ALWAYS_RNG 1      // draw the turbulence numbers of a step unconditionally: with ~24 lanes per warp some lane needs them
#endif                             // anyway, and without the branch the step is one basic block for the instruction scheduler
#ifndef SWE_TILE_PHILOX_ROUNDS
#define SWE_TILE_PHILOX_ROUNDS 10  // rounds of the counter-based generator behind the turbulence numbers (7 was measured on B200:
#endif                             // 168 against 170 ms per headline call -- not worth leaving the standard variant)
#ifndef SWE_TILE_STORE_ALL
#define SWE_TILE_STORE_ALL 1       // inner tiles store all eight neighbours of a block back (plain stores instead of eight branches)
#endif
constexpr unsigned DEAD = 0x0eu;     // pcol value of a particle that died in this super-round
constexpr unsigned PARKED = 0x0fu;   // pcol value of a particle that is not served (parked, not loaded)

// Tile geometries (host picks per call: tile_pick). 60 x 60 cells / 128 threads for maps that would not fill the GPU
// with bigger tiles; 14,400-cell tiles / 256 threads for big maps (more particles per micro-round and CTA: fuller
// warps; measured on the 4096^2 headline: 170 ms against 200 ms with 60 x 60) in three aspect ratios: 120 x 120, and
// 240 x 60 / 60 x 240 ELONGATED ALONG THE PREVAILING WIND -- a particle travels up to ~2 cells per round downwind, so a
// tile that is long in that direction parks fewer particles on its border ring (tools/par_emul.c: exported-sand bias
// against the sequential reference +2.2 % instead of +3.4 %, erosion-field correlation 0.968 instead of 0.963).
struct TileShape { int ex, ez, threads; };     // rows (x), columns (z), threads per CTA
static const TileShape kShapes[4] = {{60, 60, 128}, {120, 120, 256}, {240, 60, 256}, {60, 240, 256}};
int tile_shapes() { return 4; }
int tile_edge_x(int shape) { return kShapes[shape].ex; }
int tile_edge_z(int shape) { return kShapes[shape].ez; }
int tile_threads(int shape) { return kShapes[shape].threads; }


// one atomic per warp and destination tile
__device__ __forceinline__ void count_dest(unsigned* count, int dest, bool active) {
  const unsigned has = __ballot_sync(0xffffffffu, active);
  if (active) {
    const unsigned grp = __match_any_sync(has, dest);
    if ((int)(threadIdx.x & 31u) == __ffs(grp) - 1) atomicAdd(count + dest, (unsigned)__popc(grp));
  }
}

// colour under which the particle standing on (ix, iz) is served by the tile with origin (x0, z0), or PARKED
template <int TX, int TZ>
__device__ __forceinline__ unsigned service_colour(int ix, int iz, int x0, int z0, int nx, int nz) {
  const int lx = ix - x0, lz = iz - z0;
  if ((unsigned)lx >= (unsigned)TX || (unsigned)lz >= (unsigned)TZ) return PARKED;          // left the tile
  if ((lx == 0 && ix != 0) || (lx == TX - 1 && ix != nx - 1) || (lz == 0 && iz != 0) || (lz == TZ - 1 && iz != nz - 1))
    return PARKED;                                                                         // border ring
  return (unsigned)((lx % 3) * 3 + lz % 3);
}

struct TileGeom {
  int ex, ez;                        // tile edges in x (rows) and z (columns), multiples of 3
  int offx_cur, offz_cur, ntx_cur, ntz_cur;      // tiling served now: tile (a, b) covers [a ex - offx, (a + 1) ex - offx) x [b ez - offz, ...)
  int offx_next, offz_next, ntx_next, ntz_next;  // tiling of the next super-round
  int cur;                           // index of the record arrays / base array that hold the current segments
  int inj_lo, inj_hi;                // particle ids injected for the NEXT super-round
  unsigned micro0;                   // number of the super-round's first micro-round within the call (RNG counter)
  int K;                             // rounds per super-round
  int pcap;                          // particles of a tile that take part (the rest of its segment stays parked); multiple of 64
};

// ---- multi-GPU row stripes --------------------------------------------------------------------------------
// The tiling is GLOBAL (the same tiles as on one GPU); a tile belongs to the rank that owns the row of its centre, so
// tiles partition the map exactly as before and a super-round still needs no communication. A tile that straddles a
// stripe boundary reads and writes the neighbour's rows (heights, track vectors) directly in the neighbour's memory
// (peer mappings over NVLink: once per super-round for the heights); survivors whose next tile belongs to a neighbour
// are counted and scattered into the NEIGHBOUR's segment tables and record arrays (system-scope atomics, peer
// stores). Ranks meet three times per super-round (after fly + count, after the scan, after the scatter:
// tile_xbar_kernel) instead of once per micro-round.
struct TileStripe {
  int enabled;
  int own_lo, own_hi;                 // global rows this rank owns
  int a_lo_cur;                       // first tile row of the CURRENT tiling this rank owns
  int a_lo_next, a_hi_next;           // tile rows of the NEXT tiling this rank owns: [a_lo_next, a_hi_next)
  int a_lo_next_below;                // first tile row of the next tiling owned by the rank below (its last is a_lo_next - 1)
  int has_below, has_above;
  float* below_h; float4* below_t; int below_x0;      // neighbour's height / track plane and the global row of its first stored row
  float* above_h; float4* above_t; int above_x0;
  unsigned* below_count; unsigned* above_count;       // neighbour's per-tile counts (next tiling)
  const unsigned* below_base; const unsigned* above_base;   // neighbour's segment bases of the next tiling
  unsigned* below_cursor; unsigned* above_cursor;
  float4* below_pa; float4* below_pb; float4* above_pa; float4* above_pb;   // neighbour's record arrays of the next tiling
  unsigned* below_info; unsigned* above_info;
  int below_cap, above_cap;
};

__device__ __forceinline__ float ld_sys_f32(const float* a) {
  float v; asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(a) : "memory"); return v;
}
__device__ __forceinline__ void st_sys_f32(float* a, float v) {
  asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(a), "f"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_sys_u32(const unsigned* a) {
  unsigned v; asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(a) : "memory"); return v;
}
__device__ __forceinline__ void red_add_v4_sys(float4* addr, float a, float b, float c, float d) {
  asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// owner of tile row a of the next tiling: -1 = the rank below, 0 = this rank, +1 = the rank above
__device__ __forceinline__ int next_owner(const TileStripe& X, int a) {
  return a < X.a_lo_next ? -1 : (a >= X.a_hi_next ? 1 : 0);
}
// one atomic per warp and (owner, tile): the count lives on the owner
__device__ __forceinline__ void count_dest_stripe(const TileSched& S, const TileStripe& X, int a, int b, int ntz, bool active) {
  const int own = active ? next_owner(X, a) : 0;
  const int idx = active ? (a - (own < 0 ? X.a_lo_next_below : (own > 0 ? X.a_hi_next : X.a_lo_next))) * ntz + b : 0;
  const unsigned has = __ballot_sync(0xffffffffu, active);
  if (active) {
    const unsigned grp = __match_any_sync(has, (own + 1) * 0x10000000 + idx);
    if ((int)(threadIdx.x & 31u) == __ffs(grp) - 1) {
      if (own == 0) atomicAdd(S.count + idx, (unsigned)__popc(grp));
      else atomicAdd_system((own < 0 ? X.below_count : X.above_count) + idx, (unsigned)__popc(grp));
    }
  }
}

// Register budget: three resident 256-thread CTAs per SM (shared memory allows three 120 x 120 tiles) = at most 85
// registers. Measured on B200: left to itself with a 255-register budget ptxas takes 87, only two CTAs fit, and the
// headline call goes from 166 to 201 ms.
template <int TX, int TZ, int TNT, bool LAYERED, bool STRIPE>
__global__ void __launch_bounds__(TNT, TNT >= 256 ? 3 : 6)
tile_fly_kernel(MapView m, ParConst C, block::Thr T, TileSched S, TileGeom G, uint2 key, uint32_t call, DevStats* stats,
                const TileStripe X) {
  constexpr int TLATZ = TZ / 3, TCELLS = TX * TZ, TLAT2 = (TX / 3) * (TZ / 3);
  static_assert(TX % 3 == 0 && TZ % 3 == 0 && TNT % 32 == 0, "tile edges: multiples of 3");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = (int)threadIdx.x;
  const unsigned lane = (unsigned)tid & 31u;
  const int ntiles = G.ntx_cur * G.ntz_cur;
  const int nx = m.nx, nz = m.nz;

  if ((int)blockIdx.x >= ntiles) {
    // ---- extra CTAs: count the particles injected for the next super-round per tile of its tiling ----
    const int per_block = TNT * 8;
    const int lo = G.inj_lo + ((int)blockIdx.x - ntiles) * per_block;
#pragma unroll 1
    for (int j = 0; j < 8; j++) {
      const int id = lo + j * TNT + tid;
      bool active = id < G.inj_hi;
      int dest = 0;
      if (active) {
        float px, pz;
        spawn_position_iid(id, nx, nz, key, call, px, pz);
        const int ix = __float2int_rd(px + 0.5f), iz = __float2int_rd(pz + 0.5f);
        active = !m.oob(ix, iz);                                     // (dies in its first fly() call: never stored)
        if (active) {
          if (STRIPE) {                                              // every rank walks the whole list and keeps its own tiles' particles
            const int a = (ix + G.offx_next) / TX;
            active = a >= X.a_lo_next && a < X.a_hi_next;
            dest = (a - X.a_lo_next) * G.ntz_next + (iz + G.offz_next) / TZ;
          } else dest = ((ix + G.offx_next) / TX) * G.ntz_next + (iz + G.offz_next) / TZ;
        }
      }
      count_dest(S.count, dest, active);
    }
    return;
  }

  const unsigned* base_cur = G.cur ? S.base[1] : S.base[0];
  const unsigned seg = base_cur[blockIdx.x];
  const int n = (int)(base_cur[blockIdx.x + 1] - seg);
  if (n == 0) return;
  const int np = min(n, G.pcap);
  float4* ga = (G.cur ? S.pa[1] : S.pa[0]) + seg;      // this tile's records: nobody else touches them in this launch
  float4* gb = (G.cur ? S.pb[1] : S.pb[0]) + seg;
  const int x0 = ((STRIPE ? X.a_lo_cur : 0) + (int)blockIdx.x / G.ntz_cur) * TX - G.offx_cur, z0 = ((int)blockIdx.x % G.ntz_cur) * TZ - G.offz_cur;
  // a tile that does not touch the map's edge: every neighbour of a served cell exists
  const bool inner = x0 > 0 && z0 > 0 && x0 + TX < nx && z0 + TZ < nz;

  // ---- shared memory: heights [rock] | claim lattice | list | pcol ----
  float* hs = reinterpret_cast<float*>(smem_raw);
  float* rs = hs + (LAYERED ? TCELLS : 0);
  unsigned* claim = reinterpret_cast<unsigned*>(rs + TCELLS);
  unsigned short* list = reinterpret_cast<unsigned short*>(claim + TLAT2);
  uint4* pcol128 = reinterpret_cast<uint4*>(list + G.pcap);               // pcap is a multiple of 64
  unsigned char* pcol = reinterpret_cast<unsigned char*>(pcol128);
  __shared__ unsigned s_n[2];
  __shared__ unsigned s_stat[3];
  __shared__ float s_lost;

  if (tid < 2) s_n[tid] = 0u;
  if (tid == 2) s_lost = 0.0f;
  if (tid < 3) s_stat[tid] = 0u;
  for (int i = tid; i < TCELLS; i += TNT) {
    const int gx = x0 + i / TZ, gz = z0 + i % TZ;
    const bool in = gx >= 0 && gx < nx && gz >= 0 && gz < nz;
    if (STRIPE && in && (gx < X.own_lo || gx >= X.own_hi)) {        // a neighbour's row: read it where it lives
      hs[i] = gx < X.own_lo ? ld_sys_f32(X.below_h + (size_t)(gx - X.below_x0) * (size_t)nz + gz)
                            : ld_sys_f32(X.above_h + (size_t)(gx - X.above_x0) * (size_t)nz + gz);
      continue;
    }
    hs[i] = in ? __ldcg(m.height + (size_t)(gx - m.x0) * (size_t)nz + gz) : 0.0f;
    if (LAYERED) rs[i] = in ? __ldcg(m.rock + (size_t)(gx - m.x0) * (size_t)nz + gz) : 0.0f;
  }
  for (int i = tid; i < TLAT2; i += TNT) claim[i] = 0u;
  const int npad = (np + 15) & ~15;
  for (int p = tid; p < npad; p += TNT) {
    unsigned col = PARKED;
    if (p < np) {
      const float4 a = __ldcg(ga + p);
      col = service_colour<TX, TZ>(__float2int_rd(a.x + 0.5f), __float2int_rd(a.z + 0.5f), x0, z0, nx, nz);
    }
    pcol[p] = (unsigned char)col;
  }
  __syncthreads();

  unsigned n_steps = 0, n_oob = 0, n_age = 0;
  float lost = 0.0f;
  unsigned tag = 0u;
  int dirty = 0, dirty_rock = 0;          // this thread changed a height / a bedrock height
  const int nvec = npad >> 4;                        // 16 particles per 128-bit word of colour bytes
#pragma unroll 1
  for (int r = 0; r < G.K; r++) {
    unsigned served = 0u;
#pragma unroll 1
    for (unsigned c = 0; c < 9u; c++) {
      tag++;
      unsigned* cnt = &s_n[tag & 1u];
      // ---- collect the particles standing on interior cells of colour c: a thread scans 16 colour bytes, a warp
      //      reserves its entries with one atomic; warps beyond the tile's particle count skip the scan ----
      const unsigned pat = c * 0x01010101u;
      for (int w0 = 0; w0 < nvec; w0 += TNT) {
        if (w0 + (tid & ~31) >= nvec) break;         // (warp-uniform)
        const int w = w0 + tid;
        uint4 q = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
        if (w < nvec) q = pcol128[w];
        // bit b of byte j of e{0..3}: 0x01 per matching byte
        const unsigned e0 = __vcmpeq4(q.x, pat) & 0x01010101u, e1 = __vcmpeq4(q.y, pat) & 0x01010101u;
        const unsigned e2 = __vcmpeq4(q.z, pat) & 0x01010101u, e3 = __vcmpeq4(q.w, pat) & 0x01010101u;
        // 16-bit match mask: particle 16 w + i  (multiply gathers the four byte flags of a word into one nibble)
        const unsigned m16 = ((e0 * 0x10204080u) >> 28) | (((e1 * 0x10204080u) >> 28) << 4) |
                             (((e2 * 0x10204080u) >> 28) << 8) | (((e3 * 0x10204080u) >> 28) << 12);
        const unsigned k = (unsigned)__popc(m16);
        if (__ballot_sync(0xffffffffu, k != 0u) == 0u) continue;
        unsigned incl = k;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
          if ((int)lane >= o) incl += t;
        }
        unsigned wbase = 0u;
        if (lane == 31u) wbase = atomicAdd(cnt, incl);
        wbase = __shfl_sync(0xffffffffu, wbase, 31);
        unsigned slot = wbase + incl - k;
        for (unsigned mm = m16; mm != 0u; mm &= mm - 1u)
          list[slot++] = (unsigned short)(16 * w + __ffs(mm) - 1);
      }
      if (tid == 0) s_n[(tag + 1u) & 1u] = 0u;
      __syncthreads();
      const unsigned cntv = *cnt;
      served += cntv;
      for (unsigned i = (unsigned)tid; i < cntv; i += TNT) {
        const int p = (int)list[i];
        const float4 a = __ldcg(ga + p), b = __ldcg(gb + p);
        const int ix = __float2int_rd(a.x + 0.5f), iz = __float2int_rd(a.z + 0.5f);               // wind.hpp:30
        const int lx = ix - x0, lz = iz - z0;
        // one particle per cell per micro-round; the others wait for the colour's next turn
        if (atomicMax(&claim[(lx / 3) * TLATZ + lz / 3], tag) >= tag) continue;
        const int li = lx * TZ + lz;
        const bool xm = inner || ix > 0, xp = inner || ix + 1 < nx, zm = inner || iz > 0, zp = inner || iz + 1 < nz;
        const unsigned valid = inner ? 0xffu :
                               ((xm && zm ? 1u : 0u) | (xm ? 2u : 0u) | (xm && zp ? 4u : 0u) | (zm ? 8u : 0u) |
                                (zp ? 16u : 0u) | (xp && zm ? 32u : 0u) | (xp ? 64u : 0u) | (xp && zp ? 128u : 0u));
        const float hc = hs[li];
        float h[8];
#pragma unroll
        for (int k = 0; k < 8; k++)
          h[k] = ((valid >> k) & 1u) ? hs[li + block::slot_dx(k) * TZ + block::slot_dz(k)] : hc;
        float rock_c = 0.0f, rk[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (LAYERED) {
          rock_c = rs[li];
#pragma unroll
          for (int k = 0; k < 8; k++)
            if ((valid >> k) & 1u) rk[k] = rs[li + block::slot_dx(k) * TZ + block::slot_dz(k)];
        }
        // the order of the neighbours does not depend on the centre (swe_block.cuh: sort_neighbours), so the sorting
        // network is issued here, where it interleaves with the flight arithmetic instead of waiting for its result
        unsigned long long key8[8];
        if (!LAYERED) block::sort_neighbours(h, valid, hc, key8);
        PState q;
        q.px = a.x; q.py = a.y; q.pz = a.z; q.sed = a.w;
        q.sx = b.x; q.sy = b.y; q.sz = b.z; q.age = __float_as_int(b.w);
        const size_t ci = (size_t)ix * (size_t)nz + (size_t)iz;      // global cell number (RNG counter)
        float collision, tx, ty, tz, tsed;
        float dh = fly_owned<SWE_TILE_ALWAYS_RNG != 0, SWE_TILE_PHILOX_ROUNDS>(C, q, hc, h[1], h[6], h[3], h[4], xm, xp, zm, zp,
                             make_uint4((uint32_t)ci, call, G.micro0 + tag - 1u, 0xAu), key, collision, tx, ty, tz, tsed);
        if (STRIPE && (ix < X.own_lo || ix >= X.own_hi))                                         // wind.hpp:93-96
          red_add_v4_sys((ix < X.own_lo ? X.below_t + (size_t)(ix - X.below_x0) * (size_t)nz : X.above_t + (size_t)(ix - X.above_x0) * (size_t)nz) + iz,
                         tsed, tx, ty, tz);
        else red_add_v4(m.track + (size_t)(ix - m.x0) * (size_t)nz + (size_t)iz, tsed, tx, ty, tz);
        if (LAYERED && dh > 0.0f) {
          // layered extension: only the sand above the bedrock is picked up; the unmet demand grinds the rock
          // at `abrasion` efficiency (rock -> suspended sediment). The cell is owned: plain updates.
          const float take = fminf(dh, fmaxf(hc - rock_c, 0.0f));
          const float ground = C.abrasion * (dh - take);
          if (ground > 0.0f) { rock_c -= ground; rs[li] = rock_c; dirty_rock = 1; }
          dh = take + ground;
        }
        float c1 = hc - dh;                                                                      // :110-112
        q.sed += dh;
        float v[8];
        int s[8];
        unsigned moved;                                                                          // :114
        if (LAYERED) moved = block::cascade_sorted<true>(c1, h, valid, T, v, s, rock_c, rk);
        else {
          float r8[8];
#pragma unroll
          for (int k = 0; k < 8; k++) { v[k] = block::key_height(key8[k]); s[k] = 15 - (int)(key8[k] & 15ull); r8[k] = 0.0f; }
          moved = block::transfer_chain<false>(c1, v, s, r8, T, rock_c);
        }
        if (moved != 0u || dh != 0.0f) { hs[li] = c1; dirty = 1; }
        // slot -> (dx + 1, dz + 1) packed four bits per slot (world.hpp:83-92): two shifts instead of the branches of
        // slot_dx / slot_dz. An inner tile stores all eight neighbours (an unmoved one gets its own value back: the block
        // is owned) -- eight plain stores instead of eight branches.
#pragma unroll
        for (int i8 = 0; i8 < 8; i8++)
          if ((SWE_TILE_STORE_ALL && inner) || ((moved >> i8) & 1u)) {
            const unsigned code = (0xa9864210u >> (4 * s[i8])) & 15u;
            hs[li + (int)(code >> 2) * TZ + (int)(code & 3u) - (TZ + 1)] = v[i8];
          }
        n_steps++;
        const int jx = __float2int_rd(q.px + 0.5f), jz = __float2int_rd(q.pz + 0.5f);
        unsigned col;
        if (m.oob(jx, jz)) { n_oob++; lost += q.sed; q.age = -1; col = DEAD; }                   // :32-38 (the next fly() call)
        else if (q.age > C.max_age) { n_age++; lost += q.sed; q.age = -1; col = DEAD; }          // :45-46
        else col = service_colour<TX, TZ>(jx, jz, x0, z0, nx, nz);
        __stcg(ga + p, make_float4(q.px, q.py, q.pz, q.sed));
        __stcg(gb + p, make_float4(q.sx, q.sy, q.sz, __int_as_float(q.age)));
        pcol[p] = (unsigned char)col;
      }
      __syncthreads();
    }
    if (served == 0u) break;          // everything left in the tile is parked or dead (uniform over the CTA)
  }

  // ---- store: heights if anything moved; destination counts of the survivors (dead records carry age < 0) ----
  if (__syncthreads_or(dirty)) {
    for (int i = tid; i < TCELLS; i += TNT) {
      const int gx = x0 + i / TZ, gz = z0 + i % TZ;
      if (!(gx >= 0 && gx < nx && gz >= 0 && gz < nz)) continue;
      if (STRIPE && (gx < X.own_lo || gx >= X.own_hi))
        st_sys_f32((gx < X.own_lo ? X.below_h + (size_t)(gx - X.below_x0) * (size_t)nz : X.above_h + (size_t)(gx - X.above_x0) * (size_t)nz) + gz, hs[i]);
      else m.height[(size_t)(gx - m.x0) * (size_t)nz + gz] = hs[i];
    }
  }
  if (LAYERED && __syncthreads_or(dirty_rock)) {
    for (int i = tid; i < TCELLS; i += TNT) {
      const int gx = x0 + i / TZ, gz = z0 + i % TZ;
      if (gx >= 0 && gx < nx && gz >= 0 && gz < nz) m.rock[(size_t)(gx - m.x0) * (size_t)nz + gz] = rs[i];
    }
  }
  for (int p0 = 0; p0 < n; p0 += TNT) {
    const int p = p0 + tid;
    bool alive = false;
    int dest = 0, da = 0, db = 0;
    if (p < n) {
      alive = p >= np || pcol[p] != DEAD;              // (beyond pcap: parked for this super-round)
      if (alive) {
        const float4 a = __ldcg(ga + p);
        da = (__float2int_rd(a.x + 0.5f) + G.offx_next) / TX; db = (__float2int_rd(a.z + 0.5f) + G.offz_next) / TZ;
        dest = da * G.ntz_next + db;
      }
    }
    if (STRIPE) count_dest_stripe(S, X, da, db, G.ntz_next, alive);
    else count_dest(S.count, dest, alive);
  }

  for (int o = 16; o > 0; o >>= 1) {
    n_steps += __shfl_xor_sync(0xffffffffu, n_steps, o);
    n_oob += __shfl_xor_sync(0xffffffffu, n_oob, o);
    n_age += __shfl_xor_sync(0xffffffffu, n_age, o);
    lost += __shfl_xor_sync(0xffffffffu, lost, o);
  }
  if (lane == 0u) {
    if (n_steps) atomicAdd(&s_stat[0], n_steps);
    if (n_oob) atomicAdd(&s_stat[1], n_oob);
    if (n_age) atomicAdd(&s_stat[2], n_age);
    if (lost != 0.0f) atomicAdd(&s_lost, lost);
  }
  __syncthreads();
  if (tid == 0) {
    if (s_stat[0]) atomicAdd(&stats->steps, (unsigned long long)s_stat[0]);
    if (s_stat[1]) atomicAdd(&stats->exits_oob, (unsigned long long)s_stat[1]);
    if (s_stat[2]) atomicAdd(&stats->exits_age, (unsigned long long)s_stat[2]);
    if (s_lost != 0.0f) atomicAdd(&stats->sediment_lost, (double)s_lost);
  }
}

// Exclusive prefix sum of the per-tile counts (one CTA; a thread sums a contiguous run, the CTA scans the 1024
// partial sums); clears the counts and the scatter cursors for the next super-round.
__global__ void __launch_bounds__(1024) tile_scan_kernel(TileSched S, int nt, int nxt) {
  __shared__ unsigned part[1024];
  const int tid = (int)threadIdx.x;
  const int run = (nt + 1023) / 1024;
  const int lo = min(tid * run, nt), hi = min(lo + run, nt);
  unsigned sum = 0u;
  for (int i = lo; i < hi; i++) sum += S.count[i];
  part[tid] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const unsigned v = tid >= o ? part[tid - o] : 0u;
    __syncthreads();
    part[tid] += v;
    __syncthreads();
  }
  unsigned acc = part[tid] - sum;
  unsigned* base = nxt ? S.base[1] : S.base[0];
  for (int i = lo; i < hi; i++) {
    base[i] = acc;
    acc += S.count[i];
    S.count[i] = 0u;
    S.cursor[i] = 0u;
  }
  if (tid == 1023) { base[nt] = part[1023]; S.info[0] = part[1023]; }
}

// Moves every surviving record of the current segments, and the particles injected for the next super-round,
// into the segment of the tile (next tiling) its cell belongs to.
__global__ void __launch_bounds__(256) tile_scatter_kernel(MapView m, TileSched S, TileGeom G, uint2 key, uint32_t call,
                                                           DevStats* stats, const TileStripe X) {
  const unsigned total = (G.cur ? S.base[1] : S.base[0])[G.ntx_cur * G.ntz_cur];
  const unsigned ninj = (unsigned)(G.inj_hi - G.inj_lo);
  const unsigned all = total + ninj;
  const float4* ga = G.cur ? S.pa[1] : S.pa[0];
  const float4* gb = G.cur ? S.pb[1] : S.pb[0];
  float4* da = G.cur ? S.pa[0] : S.pa[1];
  float4* db = G.cur ? S.pb[0] : S.pb[1];
  const unsigned* base = G.cur ? S.base[0] : S.base[1];
  const unsigned lane = threadIdx.x & 31u;
  unsigned n_new = 0, n_oob = 0;
  for (unsigned j0 = blockIdx.x * 256u; j0 < all; j0 += gridDim.x * 256u) {
    const unsigned j = j0 + threadIdx.x;
    bool active = false;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < total) {
      a = __ldcs(ga + j); b = __ldcs(gb + j);
      active = __float_as_int(b.w) >= 0;
    } else if (j < all) {
      float px, pz;                                                  // Wind ctor + spawn (world.hpp:52-54)
      spawn_position_iid(G.inj_lo + (int)(j - total), m.nx, m.nz, key, call, px, pz);
      a = make_float4(px, 0.0f, pz, 0.0f);
      b = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(0));
      n_new++;
      active = !m.oob(__float2int_rd(px + 0.5f), __float2int_rd(pz + 0.5f));
      if (!active) {                                                 // dies in its first fly() call (stripes: counted by the rank whose rows are nearest)
        const int sx = __float2int_rd(px + 0.5f);
        if (!X.enabled || (sx < X.own_hi && (sx >= X.own_lo || !X.has_below)) || (sx >= X.own_hi && !X.has_above)) n_oob++;
        else n_new--;
      }
    }
    int dest = 0, own = 0;
    if (active) {
      const int ta = (__float2int_rd(a.x + 0.5f) + G.offx_next) / G.ex, tb = (__float2int_rd(a.z + 0.5f) + G.offz_next) / G.ez;
      if (X.enabled) {
        own = next_owner(X, ta);
        if (j >= total) {                              // a new particle: every rank walks the whole list and keeps its own
          if (own != 0) { active = false; n_new--; }
        }
        dest = (ta - (own < 0 ? X.a_lo_next_below : (own > 0 ? X.a_hi_next : X.a_lo_next))) * G.ntz_next + tb;
        if (active && ((own < 0 && !X.has_below) || (own > 0 && !X.has_above) || dest < 0)) {   // cannot happen: stripes are taller than a tile
          active = false; atomicAdd(S.info + 1, 1u);
        }
      } else dest = ta * G.ntz_next + tb;
    }
    const unsigned has = __ballot_sync(0xffffffffu, active);
    if (active) {
      const unsigned grp = __match_any_sync(has, (own + 1) * 0x10000000 + dest);
      const int leader = __ffs(grp) - 1;
      unsigned wb = 0u;
      if ((int)lane == leader)
        wb = own == 0 ? atomicAdd(S.cursor + dest, (unsigned)__popc(grp))
                      : atomicAdd_system((own < 0 ? X.below_cursor : X.above_cursor) + dest, (unsigned)__popc(grp));
      wb = __shfl_sync(grp, wb, leader);
      const unsigned b0 = own == 0 ? base[dest] : ld_sys_u32((own < 0 ? X.below_base : X.above_base) + dest);
      const unsigned slot = b0 + wb + (unsigned)__popc(grp & ((1u << lane) - 1u));
      float4* oa = own == 0 ? da : (own < 0 ? X.below_pa : X.above_pa);
      float4* ob = own == 0 ? db : (own < 0 ? X.below_pb : X.above_pb);
      const unsigned ocap = own == 0 ? (unsigned)S.cap : (unsigned)(own < 0 ? X.below_cap : X.above_cap);
      if (slot < ocap) { oa[slot] = a; ob[slot] = b; }
      else atomicAdd(S.info + 1, 1u);                                // cannot happen: cap >= particles per call
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    n_new += __shfl_xor_sync(0xffffffffu, n_new, o);
    n_oob += __shfl_xor_sync(0xffffffffu, n_oob, o);
  }
  if (lane == 0u) {
    if (n_new) atomicAdd(&stats->particles, (unsigned long long)n_new);
    if (n_oob) atomicAdd(&stats->exits_oob, (unsigned long long)n_oob);
  }
}

// Cross-GPU barrier between the phases of a super-round (one warp, thread 0 works). Every rank publishes
// (sequence number, records it holds for the next super-round) into EVERY rank's mailbox ring (16 slots x 8 ranks of
// 8-byte words, one 64-byte line per slot) and waits for its two NEIGHBOURS only -- the ranks whose memory it reads
// and writes. want_sum: also wait for all ranks' words of this sequence number (every rank runs the same sequence,
// so they all arrive) and leave the global number of records in info[2] for the host's completion check.
struct TileMail { unsigned long long* box[8]; int rank, nranks; };
__global__ void tile_xbar_kernel(TileMail M, unsigned long long seq, unsigned* info, int want_sum) {
  if (threadIdx.x != 0) return;
  asm volatile("fence.acq_rel.sys;" ::: "memory");
  const unsigned long long word = (seq << 32) | (unsigned long long)info[0];
  const int slot = (int)(seq & 15ull);
  for (int i = 0; i < M.nranks; i++)
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(M.box[i] + slot * 8 + M.rank), "l"(word) : "memory");
  const unsigned long long* mine = M.box[M.rank];
  const long long t0 = clock64();
  unsigned long long total = 0ull;
  for (int i = 0; i < M.nranks; i++) {
    const bool neighbour = i == M.rank - 1 || i == M.rank + 1;
    if (!neighbour && !want_sum) continue;
    unsigned long long w;
    unsigned spins = 0;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(mine + slot * 8 + i) : "memory");
      if ((w >> 32) == (seq & 0xffffffffull)) break;
      __nanosleep(100);
      if ((++spins & 1023u) == 0u && clock64() - t0 > 16000000000LL) { info[3] = 1u; return; }   // ~8 s: never hang the device
    }
    total += w & 0xffffffffull;
  }
  if (want_sum) info[2] = (unsigned)(total > 0xffffffffull ? 0xffffffffull : total);
  asm volatile("fence.acq_rel.sys;" ::: "memory");
}

void launch_tile_xbar(void* const* mailboxes, int rank, int nranks, unsigned long long seq, unsigned* info, bool want_sum, cudaStream_t st) {
  TileMail M{};
  for (int i = 0; i < nranks && i < 8; i++) M.box[i] = static_cast<unsigned long long*>(mailboxes[i]);
  M.rank = rank; M.nranks = nranks;
  tile_xbar_kernel<<<1, 32, 0, st>>>(M, seq, info, want_sum ? 1 : 0);
}

static size_t tile_smem_bytes(int shape, bool layered, int pcap) {
  const int ex = kShapes[shape].ex, ez = kShapes[shape].ez;
  return (size_t)(layered ? 2 : 1) * ex * ez * sizeof(float) + (size_t)pcap * (sizeof(unsigned short) + 1) +
         (size_t)(ex / 3) * (ez / 3) * sizeof(unsigned) + 16;
}

// largest pcap (multiple of 64) the device can give one CTA; the list holds 16-bit particle indices
int tile_max_pcap(int shape, bool layered, int smem_optin) {
  const long long fixed = (long long)tile_smem_bytes(shape, layered, 0) + 1024;
  long long p = ((long long)smem_optin - fixed) / 3;
  p = p / 64 * 64;
  return p < 64 ? 0 : (int)(p > 32768 ? 32768 : p);
}

static TileGeom tile_geom(const MapView& m, int ex, int ez, int s) {
  TileGeom G{};
  G.ex = ex; G.ez = ez;
  G.offx_cur = (s % 3) * (ex / 3); G.offz_cur = (s % 3) * (ez / 3);
  G.ntx_cur = (m.nx + G.offx_cur + ex - 1) / ex; G.ntz_cur = (m.nz + G.offz_cur + ez - 1) / ez;
  G.offx_next = ((s + 1) % 3) * (ex / 3); G.offz_next = ((s + 1) % 3) * (ez / 3);
  G.ntx_next = (m.nx + G.offx_next + ex - 1) / ex; G.ntz_next = (m.nz + G.offz_next + ez - 1) / ez;
  G.cur = s & 1;
  return G;
}

// tiles of the tiling with the most tiles (allocation size of the per-tile arrays: the smallest edge, the largest shift)
int tile_max_tiles(int nx, int nz) {
  int best = 0;
  for (const TileShape& sh : kShapes) {
    const int t = ((nx + 2 * (sh.ex / 3) + sh.ex - 1) / sh.ex) * ((nz + 2 * (sh.ez / 3) + sh.ez - 1) / sh.ez);
    if (t > best) best = t;
  }
  return best;
}

// Shape for a map and a prevailing wind (px, pz): the big tiles when they still give every SM a few CTAs, elongated
// along the wind when it blows mostly along one axis
int tile_pick(int nx, int nz, int sm_count, float px, float pz) {
  const long long tiles = (long long)((nx + 119) / 120) * ((nz + 119) / 120);
  if (tiles < 3LL * sm_count) return 0;
  const float ax = fabsf(px), az = fabsf(pz);
  if (ax >= 2.0f * az && nx >= 2 * kShapes[2].ex) return 2;
  if (az >= 2.0f * ax && nz >= 2 * kShapes[3].ez) return 3;
  return 1;
}

// stripes: first tile row (tiling offset off) whose centre row lies at or beyond global row x
static int first_tile_row(int x, int tt, int off, int nx) {
  const int ntx = (nx + off + tt - 1) / tt;
  if (x <= 0) return 0;
  if (x >= nx) return ntx;
  const int num = x + off - tt / 2;
  int a = num <= 0 ? 0 : (num + tt - 1) / tt;
  return a > ntx ? ntx : a;
}

// Host arithmetic of the stripe ownership, exported for tests (no device needed): the tile rows [a_lo, a_hi) of
// tiling `tiling` (0, 1, 2) that the rank owning global rows [own_lo, own_hi) launches, and the number of tile rows.
void tile_owned_rows(int nx, int shape, int tiling, int own_lo, int own_hi, int* a_lo, int* a_hi, int* ntx) {
  const int tt = kShapes[shape].ex, off = (tiling % 3) * (tt / 3);
  *a_lo = first_tile_row(own_lo, tt, off, nx);
  *a_hi = first_tile_row(own_hi, tt, off, nx);
  *ntx = (nx + off + tt - 1) / tt;
}

template <int TX, int TZ, int TNT>
static void launch_fly(bool layered, bool stripe, int grid, size_t smem, cudaStream_t st, const MapView& m, const ParConst& C,
                       const block::Thr& T, const TileSched& S, const TileGeom& G, uint2 key, uint32_t call, DevStats* stats,
                       const TileStripe& X) {
  if (stripe) tile_fly_kernel<TX, TZ, TNT, false, true><<<grid, TNT, smem, st>>>(m, C, T, S, G, key, call, stats, X);
  else if (layered) tile_fly_kernel<TX, TZ, TNT, true, false><<<grid, TNT, smem, st>>>(m, C, T, S, G, key, call, stats, X);
  else tile_fly_kernel<TX, TZ, TNT, false, false><<<grid, TNT, smem, st>>>(m, C, T, S, G, key, call, stats, X);
}

// Geometry of super-round s (s = -1: only the injection of super-round 0's particles) for a single-GPU engine or one
// rank of a stripe group.
static void tile_round_setup(const MapView& m, int shape, int s, int K, int pcap, int inj_lo, int inj_hi,
                             const TileSched& S, const TileStripeHost* H, TileGeom& G, TileStripe& X) {
  const int tt = kShapes[shape].ex;           // (stripes cut the map along x: tile rows)
  if (s >= 0) G = tile_geom(m, kShapes[shape].ex, kShapes[shape].ez, s);
  else {                                      // no tiles yet: cur = 1 holds an empty segment table (base[1][0] = 0)
    G = tile_geom(m, kShapes[shape].ex, kShapes[shape].ez, 2);   // (s + 1) % 3 == 0  -> next tiling is tiling 0
    G.ntx_cur = 0; G.ntz_cur = 0; G.cur = 1;
  }
  G.inj_lo = inj_lo; G.inj_hi = inj_hi;
  G.micro0 = s >= 0 ? (unsigned)s * 9u * (unsigned)K : 0u;
  G.K = K; G.pcap = pcap;
  X = TileStripe{};
  if (!H || !H->enabled) return;
  const int nxt = G.cur ^ 1;
  X.enabled = 1; X.own_lo = H->own_lo; X.own_hi = H->own_hi;
  X.has_below = H->has_below; X.has_above = H->has_above;
  const int a_lo_cur = first_tile_row(H->own_lo, tt, G.offx_cur, m.nx), a_hi_cur = first_tile_row(H->own_hi, tt, G.offx_cur, m.nx);
  X.a_lo_cur = a_lo_cur;
  if (s >= 0) G.ntx_cur = a_hi_cur - a_lo_cur;                       // tile rows this rank launches
  X.a_lo_next = first_tile_row(H->own_lo, tt, G.offx_next, m.nx);
  X.a_hi_next = first_tile_row(H->own_hi, tt, G.offx_next, m.nx);
  X.a_lo_next_below = H->has_below ? first_tile_row(H->below_lo, tt, G.offx_next, m.nx) : 0;
  G.ntx_next = X.a_hi_next - X.a_lo_next;                            // tile rows of the next tiling this rank owns (scan length)
  X.below_h = H->below_h; X.below_t = H->below_t; X.below_x0 = H->below_x0;
  X.above_h = H->above_h; X.above_t = H->above_t; X.above_x0 = H->above_x0;
  X.below_count = H->below.count; X.above_count = H->above.count;
  X.below_base = H->below.base[nxt]; X.above_base = H->above.base[nxt];
  X.below_cursor = H->below.cursor; X.above_cursor = H->above.cursor;
  X.below_pa = H->below.pa[nxt]; X.below_pb = H->below.pb[nxt];
  X.above_pa = H->above.pa[nxt]; X.above_pb = H->above.pb[nxt];
  X.below_info = H->below.info; X.above_info = H->above.info;
  X.below_cap = H->below.cap; X.above_cap = H->above.cap;
}

// The three phases of a super-round. phase 0: fly + count (skipped when there is nothing to launch), 1: scan, 2: scatter.
// Stripe groups put a cross-GPU barrier (launch_tile_xbar) after each of them. Returns the launches made.
int launch_tile_phase(int phase, const MapView& m, const swe_params& p, const TileSched& S, int shape, int s, int K, int pcap,
                      int inj_lo, int inj_hi, uint64_t seed, uint32_t call, DevStats* stats, cudaStream_t st,
                      const TileStripeHost* H) {
  const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  TileGeom G;
  TileStripe X;
  tile_round_setup(m, shape, s, K, pcap, inj_lo, inj_hi, S, H, G, X);
  const int nt = kShapes[shape].threads;
  if (phase == 0) {
    block::Thr T;
    T.ortho = sqrtf(1.0f) * p.maxdiff * 1.0f;   // host IEEE fp32: the values exact::cascade_thresholds computes
    T.diag = sqrtf(2.0f) * p.maxdiff * 1.0f;
    T.settling = p.settling;
    const int ntiles = G.ntx_cur * G.ntz_cur;
    const int inj_blocks = (inj_hi - inj_lo + nt * 8 - 1) / (nt * 8);
    const bool layered = m.rock != nullptr;
    const size_t smem = tile_smem_bytes(shape, layered, pcap);
    if (ntiles + inj_blocks <= 0) return 0;
    const int grid = ntiles + inj_blocks;
    const bool stripe = X.enabled != 0;
    switch (shape) {
      case 1: launch_fly<120, 120, 256>(layered, stripe, grid, smem, st, m, make_const(p), T, S, G, key, call, stats, X); break;
      case 2: launch_fly<240, 60, 256>(layered, stripe, grid, smem, st, m, make_const(p), T, S, G, key, call, stats, X); break;
      case 3: launch_fly<60, 240, 256>(layered, stripe, grid, smem, st, m, make_const(p), T, S, G, key, call, stats, X); break;
      default: launch_fly<60, 60, 128>(layered, stripe, grid, smem, st, m, make_const(p), T, S, G, key, call, stats, X); break;
    }
    return 1;
  }
  if (phase == 1) {
    tile_scan_kernel<<<1, 1024, 0, st>>>(S, G.ntx_next * G.ntz_next, G.cur ^ 1);
    return 1;
  }
  long long blocks = ((long long)S.cap + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  tile_scatter_kernel<<<(int)blocks, 256, 0, st>>>(m, S, G, key, call, stats, X);
  return 1;
}

int launch_tile_super_round(const MapView& m, const swe_params& p, const TileSched& S, int shape, int s, int K, int pcap,
                            int inj_lo, int inj_hi, uint64_t seed, uint32_t call, DevStats* stats, cudaStream_t st) {
  int n = 0;
  for (int phase = 0; phase < 3; phase++)
    n += launch_tile_phase(phase, m, p, S, shape, s, K, pcap, inj_lo, inj_hi, seed, call, stats, st, nullptr);
  return n;
}

template <int TX, int TZ, int TNT>
static cudaError_t configure_fly(bool layered, int smem) {
  const cudaFuncAttribute at = cudaFuncAttributeMaxDynamicSharedMemorySize;
  if (layered) return cudaFuncSetAttribute(tile_fly_kernel<TX, TZ, TNT, true, false>, at, smem);
  cudaError_t err = cudaFuncSetAttribute(tile_fly_kernel<TX, TZ, TNT, false, false>, at, smem);
  if (err == cudaSuccess) err = cudaFuncSetAttribute(tile_fly_kernel<TX, TZ, TNT, false, true>, at, smem);
  return err;
}

cudaError_t tile_configure(int shape, bool layered, int pcap) {
  const int smem = (int)tile_smem_bytes(shape, layered, pcap);
  switch (shape) {
    case 1: return configure_fly<120, 120, 256>(layered, smem);
    case 2: return configure_fly<240, 60, 256>(layered, smem);
    case 3: return configure_fly<60, 240, 256>(layered, smem);
    default: return configure_fly<60, 60, 128>(layered, smem);
  }
}

template <int TX, int TZ, int TNT>
static void preload_fly() {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, tile_fly_kernel<TX, TZ, TNT, false, false>);
  cudaFuncGetAttributes(&a, tile_fly_kernel<TX, TZ, TNT, true, false>);
  cudaFuncGetAttributes(&a, tile_fly_kernel<TX, TZ, TNT, false, true>);
}

void preload_tile() {
  preload_fly<60, 60, 128>(); preload_fly<120, 120, 256>(); preload_fly<240, 60, 256>(); preload_fly<60, 240, 256>();
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, tile_scan_kernel);
  cudaFuncGetAttributes(&a, tile_scatter_kernel);
  cudaFuncGetAttributes(&a, tile_xbar_kernel);
}

}  // namespace swe
