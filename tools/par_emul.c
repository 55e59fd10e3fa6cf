/* TOOL (test infrastructure, CPU): emulator of the PARALLEL-mode semantics on top of the oracle's
 * physics, used to study how the statistical parity with the sequential reference depends on the
 * scheduling of the particles (all at once / batches / staggered starts / visibility of erosion
 * within a round / where the cascade runs). Not part of the product; design input for DESIGN.md 3.
 *   gcc -O2 -ffp-contract=off -shared -fPIC tools/par_emul.c -o /tmp/libpar_emul.so -lm
 */
#include <stdio.h>
#include "../oracle/swe_oracle.c"

/* Wind::fly (wind.hpp:28-112) without the cascade; erosion goes to hd[] when given (frozen
 * visibility: applied by the caller after the round), else straight into the height plane. */
static int fly_nc(swo_world* w, swo_particle* q, float* hd, int* cell) {
  const swo_params* P = &w->p;
  const int ix = (int)roundf(q->pos[0]), iz = (int)roundf(q->pos[2]);
  if (oob(w, ix, iz)) return 0;
  const size_t ind = (size_t)ix * w->nz + iz;
  const float h = w->plane[0][ind];
  const v3 n = normal_v3(w, ix, iz);
  v3 pos = v3_(q->pos[0], q->pos[1], q->pos[2]);
  v3 speed = v3_(q->speed[0], q->speed[1], q->speed[2]);
  const v3 pspeed = v3_(P->pspeed[0], P->pspeed[1], P->pspeed[2]);
  if (q->age == 0 || pos.y < h) pos.y = h;
  q->pos[1] = pos.y;
  if (q->age++ > P->max_age) return 0;
  float hfac = swo_expf(-(pos.y - h) / P->boundarylayer);
  if (hfac < 0) hfac = 0;
  float shadow = v3_dot(v3_normalize(pspeed), n);
  if (shadow < 0) shadow = 0;
  shadow = 1.0f - shadow;
  speed = v3_add(speed, v3_scale(0.05f, v3_sub(v3_scale(0.1f + 0.9f * shadow, pspeed), speed)));
  if (pos.y > h) speed.y -= P->gravity * q->sediment;
  float collision = -v3_dot(v3_normalize(speed), n);
  if (collision < 0) collision = 0;
  v3 rspeed = v3_cross(n, v3_cross(v3_scale(1.0f - collision, speed), n));
  speed = v3_add(speed, v3_scale(0.9f, v3_sub(v3_scale(shadow, v3_mix(pspeed, rspeed, shadow * hfac)), speed)));
  {
    int rz = swo_rand(&w->rng) % 1001, ry = swo_rand(&w->rng) % 1001, rx = swo_rand(&w->rng) % 1001;
    v3 t = v3_((float)rx - 500.0f, (float)ry - 500.0f, (float)rz - 500.0f);
    t = v3_scale(0.1f * hfac * collision, t);
    t = v3_(t.x / 500.0f, t.y / 500.0f, t.z / 500.0f);
    speed = v3_add(speed, t);
  }
  {
    float f = (float)(1.0 - 0.3 * (double)q->sediment);
    speed = v3_(speed.x * f, speed.y * f, speed.z * f);
  }
  pos = v3_add(pos, speed);
  w->plane[6][ind] += speed.x; w->plane[7][ind] += speed.y; w->plane[8][ind] += speed.z;
  w->plane[5][ind] += q->sediment;
  float force = -v3_dot(v3_normalize(speed), n) * v3_length(speed);
  if (force < 0) force = 0;
  float lift = (1.0f - collision) * v3_length(speed);
  float capacity = force * hfac + 0.02f * lift * hfac;
  float diff = capacity - q->sediment;
  if (hd) hd[ind] -= P->suspension * diff; else w->plane[0][ind] -= P->suspension * diff;
  q->sediment += P->suspension * diff;
  q->pos[0] = pos.x; q->pos[1] = pos.y; q->pos[2] = pos.z;
  q->speed[0] = speed.x; q->speed[1] = speed.y; q->speed[2] = speed.z;
  *cell = (int)ind;
  return 1;
}

/* visibility 0: erosion of a round becomes visible after the round (GPU atomics semantics, idealised)
 *            1: immediately (particles of a round see each other in particle order)
 * cascade    0: World::cascade once at every touched cell after the round, nine-colour order
 *            1: World::cascade right after the particle's step (needs visibility 1)
 *            2: touched cells after the round, in particle order (no colouring)
 * stagger  S: particle i starts at round (i * S) / n  (S = 0: all at round 0)
 * K: steps per particle per round. */
long long emul_erode(swo_world* w, int cycles, int batches, int stagger, int visibility, int cascade, int K,
                     double* lost_out, long long* rounds_out) {
  const size_t area = (size_t)w->nx * w->nz;
  for (int k = 5; k < 9; k++) memset(w->plane[k], 0, sizeof(float) * area);
  long long steps = 0, rounds = 0;
  double lost = 0.0;
  float* hd = visibility != 1 ? (float*)calloc(area, sizeof(float)) : NULL;
  unsigned char* touched = (unsigned char*)calloc(area, 1);
  int* tlist = (int*)malloc(sizeof(int) * (size_t)(cycles + 1) * (size_t)(K > 0 ? K : 1));
  for (int b = 0; b < batches; b++) {
    const int n = cycles / batches + (b < cycles % batches ? 1 : 0);
    swo_particle* q = (swo_particle*)malloc(sizeof(swo_particle) * (size_t)(n + 1));
    int* start = (int*)malloc(sizeof(int) * (size_t)(n + 1));
    char* alive = (char*)malloc((size_t)n + 1);
    for (int i = 0; i < n; i++) { swo_spawn(w, &q[i]); alive[i] = 1; start[i] = stagger > 0 ? (int)(((long long)i * stagger) / n) : 0; }
    /* staggered starts: shuffle start rounds so that they are independent of the spawn order */
    int live = n;
    for (int r = 0; live > 0; r++, rounds++) {
      int nt = 0;
      for (int i = 0; i < n; i++) {
        if (!alive[i] || start[i] > r) continue;
        if (visibility == 2) {
          const int ix = (int)roundf(q[i].pos[0]), iz = (int)roundf(q[i].pos[2]);
          if (!oob(w, ix, iz) && touched[(size_t)ix * w->nz + iz]) continue;   /* cell taken this round */
        }
        for (int s = 0; s < K; s++) {
          int cell;
          if (!fly_nc(w, &q[i], hd, &cell)) { alive[i] = 0; live--; lost += (double)q[i].sediment; break; }
          steps++;
          if (cascade == 1) swo_cascade(w, cell / w->nz, cell % w->nz);
          else if (!touched[cell]) { touched[cell] = 1; tlist[nt++] = cell; }
          else if (cascade == 2) tlist[nt++] = cell;
        }
      }
      if (hd) for (int t = 0; t < nt; t++) { w->plane[0][tlist[t]] += hd[tlist[t]]; hd[tlist[t]] = 0.0f; }
      if (cascade == 0) {
        for (int c = 0; c < 9; c++)
          for (int t = 0; t < nt; t++) {
            const int x = tlist[t] / w->nz, z = tlist[t] % w->nz;
            if (x % 3 == c / 3 && z % 3 == c % 3) swo_cascade(w, x, z);
          }
      } else if (cascade == 2) {
        for (int t = 0; t < nt; t++) swo_cascade(w, tlist[t] / w->nz, tlist[t] % w->nz);
      }
      for (int t = 0; t < nt; t++) touched[tlist[t]] = 0;
    }
    free(q); free(start); free(alive);
  }
  const float lr = w->p.lrate;
  for (int k = 1; k < 5; k++) {
    float* f = w->plane[k]; const float* t = w->plane[k + 4];
    for (size_t i = 0; i < area; i++) f[i] = (1.0f - lr) * f[i] + lr * t[i];
  }
  free(hd); free(touched); free(tlist);
  if (lost_out) *lost_out = lost;
  if (rounds_out) *rounds_out = rounds;
  return steps;
}

/* Mode 3: colour-phased rounds. In every round the particles are processed in nine colour phases by the
 * colour of their CURRENT cell; within a phase a particle flies and cascades at once (reference order of
 * operations, wind.hpp:110-114); a particle that finds its cell already claimed in this round skips the
 * round (claim_skip != 0) or is simply serialised (claim_skip == 0).
 * Mode 4: the same, but every T x T tile works on a private copy of its (T+2)^2 window taken at the start
 * of the round and the height differences are added to the map afterwards (what per-tile shared-memory
 * CTAs do): erosion and cascade transfers across tile borders are stale within a round. */
long long emul_erode_phased(swo_world* w, int cycles, int batches, int claim_skip, int T, int stagger, double* lost_out,
                            long long* rounds_out) {
  const int nx = w->nx, nz = w->nz;
  const size_t area = (size_t)nx * nz;
  for (int k = 5; k < 9; k++) memset(w->plane[k], 0, sizeof(float) * area);
  long long steps = 0, rounds = 0;
  double lost = 0.0;
  int* claim = (int*)calloc(area, sizeof(int));
  float* G = w->plane[0];
  float* snap = T > 0 ? (float*)malloc(sizeof(float) * area) : NULL;
  float* loc = T > 0 ? (float*)malloc(sizeof(float) * area) : NULL;   /* all tiles' private copies: interior part */
  int stamp = 0;
  for (int b = 0; b < batches; b++) {
    const int n = cycles / batches + (b < cycles % batches ? 1 : 0);
    swo_particle* q = (swo_particle*)malloc(sizeof(swo_particle) * (size_t)(n + 1));
    char* alive = (char*)malloc((size_t)n + 1);
    int* col = (int*)malloc(sizeof(int) * (size_t)(n + 1));
    for (int i = 0; i < n; i++) { swo_spawn(w, &q[i]); alive[i] = 1; }
    int live = n;
    long long r0 = rounds;
    while (live > 0) {
      rounds++; stamp++;
      for (int i = 0; i < n; i++) {
        col[i] = -1;
        if (!alive[i]) continue;
        if (stagger > 0 && ((long long)i * stagger) / n > rounds - 1 - r0) continue;
        const int ix = (int)roundf(q[i].pos[0]), iz = (int)roundf(q[i].pos[2]);
        if (oob(w, ix, iz)) { alive[i] = 0; live--; lost += (double)q[i].sediment; col[i] = -1; continue; }
        col[i] = (ix % 3) * 3 + iz % 3;
      }
      if (T <= 0 && (claim_skip & 2)) {
        /* frozen fly reads: every claim winner flies against the heights of the round start (erosion goes
         * to a side buffer), then erosion + World::cascade are applied winner by winner in colour order */
        static float* hd = NULL; static size_t hd_n = 0;
        if (hd_n != area) { free(hd); hd = (float*)calloc(area, sizeof(float)); hd_n = area; }
        for (int i = 0; i < n; i++) {
          if (!alive[i] || col[i] < 0) continue;
          const int ix = (int)roundf(q[i].pos[0]), iz = (int)roundf(q[i].pos[2]);
          const size_t ind = (size_t)ix * nz + iz;
          if ((claim_skip & 1) && claim[ind] == stamp) { col[i] = -1; continue; }
          claim[ind] = stamp;
          int cell;
          if (!fly_nc(w, &q[i], hd, &cell)) { alive[i] = 0; live--; lost += (double)q[i].sediment; col[i] = -1; continue; }
          steps++;
          col[i] = 16 + col[i];   /* flew: remember the colour of the cell it eroded */
          q[i].age |= 0;          /* (cell recomputed below from the claim order) */
          /* stash the eroded cell index in start-of-round order: reuse claim via negative encoding not needed */
        }
        /* apply in colour order: cells with hd != 0 or claimed this round */
        for (int c = 0; c < 9; c++)
          for (int x = c / 3; x < nx; x += 3)
            for (int z = c % 3; z < nz; z += 3) {
              const size_t ind = (size_t)x * nz + z;
              if (claim[ind] != stamp) continue;
              G[ind] += hd[ind]; hd[ind] = 0.0f;
              swo_cascade(w, x, z);
            }
      } else if (T <= 0) {
        for (int c = 0; c < 9; c++)
          for (int i = 0; i < n; i++) {
            if (!alive[i] || col[i] != c) continue;
            const int ix = (int)roundf(q[i].pos[0]), iz = (int)roundf(q[i].pos[2]);
            const size_t ind = (size_t)ix * nz + iz;
            if (claim_skip && claim[ind] == stamp) continue;
            claim[ind] = stamp;
            int cell;
            if (!fly_nc(w, &q[i], NULL, &cell)) { alive[i] = 0; live--; lost += (double)q[i].sediment; continue; }
            steps++;
            swo_cascade(w, ix, iz);
          }
      } else {
        /* tile-local: every tile sees the snapshot + its own changes only */
        memcpy(snap, G, sizeof(float) * area);
        float* acc = (float*)calloc(area, sizeof(float));      /* sum of all tiles' deltas */
        float* win = (float*)malloc(sizeof(float) * (size_t)(T + 2) * (T + 2));
        for (int tx = 0; tx < nx; tx += T)
          for (int tz = 0; tz < nz; tz += T) {
            /* private window = snapshot; run this tile's particles on the MAP itself after swapping the window in */
            const int x0 = tx - 1 < 0 ? 0 : tx - 1, x1 = tx + T + 1 > nx ? nx : tx + T + 1;
            const int z0 = tz - 1 < 0 ? 0 : tz - 1, z1 = tz + T + 1 > nz ? nz : tz + T + 1;
            for (int x = x0; x < x1; x++) for (int z = z0; z < z1; z++) G[(size_t)x * nz + z] = snap[(size_t)x * nz + z];
            for (int c = 0; c < 9; c++)
              for (int i = 0; i < n; i++) {
                if (!alive[i] || col[i] != c) continue;
                const int ix = (int)roundf(q[i].pos[0]), iz = (int)roundf(q[i].pos[2]);
                if (ix < tx || ix >= tx + T || iz < tz || iz >= tz + T) continue;
                col[i] = -2;   /* stepped this round */
                int cell;
                if (!fly_nc(w, &q[i], NULL, &cell)) { alive[i] = 0; live--; lost += (double)q[i].sediment; continue; }
                steps++;
                swo_cascade(w, ix, iz);
              }
            for (int x = x0; x < x1; x++) for (int z = z0; z < z1; z++) {
              const size_t j = (size_t)x * nz + z;
              acc[j] += G[j] - snap[j];
              G[j] = snap[j];
            }
          }
        for (size_t j = 0; j < area; j++) G[j] = snap[j] + acc[j];
        free(acc); free(win);
      }
    }
    free(q); free(alive); free(col);
  }
  const float lr = w->p.lrate;
  for (int k = 1; k < 5; k++) {
    float* f = w->plane[k]; const float* t = w->plane[k + 4];
    for (size_t i = 0; i < area; i++) f[i] = (1.0f - lr) * f[i] + lr * t[i];
  }
  free(claim); free(snap); free(loc);
  if (lost_out) *lost_out = lost;
  if (rounds_out) *rounds_out = rounds;
  return steps;
}

/* Option C: colour-queue micro-rounds. Micro-round k serves the colour order[k % 9]: every particle standing
 * on a cell of that colour tries to claim its cell (one winner per cell), flies one step and runs
 * World::cascade at once (3x3 blocks of one colour are disjoint, so the micro-round is conflict-free and
 * equivalent to processing its particles one after another). Particle i is injected at micro-round
 * i * stagger / n. */
long long emul_erode_queues(swo_world* w, int cycles, int stagger, int order_kind, double* lost_out,
                            long long* rounds_out) {
  const int nx = w->nx, nz = w->nz;
  const size_t area = (size_t)nx * nz;
  for (int k = 5; k < 9; k++) memset(w->plane[k], 0, sizeof(float) * area);
  long long steps = 0, k = 0;
  double lost = 0.0;
  int* claim = (int*)calloc(area, sizeof(int));
  const int n = cycles;
  swo_particle* q = (swo_particle*)malloc(sizeof(swo_particle) * (size_t)(n + 1));
  char* alive = (char*)malloc((size_t)n + 1);
  for (int i = 0; i < n; i++) { swo_spawn(w, &q[i]); alive[i] = 1; }
  int live = n;
  static const int ord0[9] = {0, 1, 2, 3, 4, 5, 6, 7, 8};      /* z fastest */
  static const int ord1[9] = {0, 3, 6, 1, 4, 7, 2, 5, 8};      /* x fastest: a +x mover is served again at once */
  const int* order = (order_kind & 1) ? ord1 : ord0;
  while (live > 0) {
    const int c = order[k % 9];
    const int stamp = (int)(k + 1);
    for (int i = 0; i < n; i++) {
      if (!alive[i]) continue;
      if (stagger > 0 && ((long long)i * stagger) / n > k) continue;
      const int ix = (int)roundf(q[i].pos[0]), iz = (int)roundf(q[i].pos[2]);
      if (oob(w, ix, iz)) { alive[i] = 0; live--; lost += (double)q[i].sediment; continue; }
      if ((ix % 3) * 3 + iz % 3 != c) continue;
      const size_t ind = (size_t)ix * nz + iz;
      if (claim[ind] == stamp) continue;
      claim[ind] = stamp;
      /* order_kind >= 16: a particle that stays on its cell keeps its claim and steps again at once, up to
       * (order_kind >> 4) steps per micro-round (it owns the cell and its 3x3 block for the whole micro-round) */
      const int reps = order_kind >= 16 ? (order_kind >> 4) : 1;
      for (int rep = 0; rep < reps; rep++) {
        int cell;
        if (!fly_nc(w, &q[i], NULL, &cell)) { alive[i] = 0; live--; lost += (double)q[i].sediment; break; }
        steps++;
        swo_cascade(w, ix, iz);
        if ((int)roundf(q[i].pos[0]) != ix || (int)roundf(q[i].pos[2]) != iz) break;
      }
    }
    k++;
  }
  const float lr = w->p.lrate;
  for (int kk = 1; kk < 5; kk++) {
    float* f = w->plane[kk]; const float* t = w->plane[kk + 4];
    for (size_t i = 0; i < area; i++) f[i] = (1.0f - lr) * f[i] + lr * t[i];
  }
  free(claim); free(q); free(alive);
  if (lost_out) *lost_out = lost;
  if (rounds_out) *rounds_out = k;
  return steps;
}

/* Option D: TILE-RESIDENT super-rounds (what csrc/swe_tile.cu does). The map is cut into T x T tiles; a
 * super-round runs K rounds of nine colour micro-rounds INSIDE every tile independently (one CTA per tile,
 * heights and particles in shared memory). A particle is served only while its cell lies in the INTERIOR of
 * the tile it started the super-round in (>= 1 cell from the tile's edges, map edges excepted), so that its
 * 3x3 block never leaves the tile: tiles are conflict-free, no staleness, no atomics on heights. A particle
 * that reaches the border ring (or jumps out of the tile) is parked until the next super-round, whose tiling
 * is shifted by T/3 in both directions (three tilings in turn: a cell lies on the x-ring of at most one of
 * them and on the z-ring of at most one, so it is interior in at least one).
 * Particle i becomes active at round i * stagger / n (flags & 1: at that round inside the super-round;
 * else at the start of the super-round that contains it). */
long long emul_erode_tiles(swo_world* w, int cycles, int stagger, int T, int K, int flags, double* lost_out,
                           long long* rounds_out) {
  const int nx = w->nx, nz = w->nz;
  const size_t area = (size_t)nx * nz;
  for (int k = 5; k < 9; k++) memset(w->plane[k], 0, sizeof(float) * area);
  long long steps = 0, sr = 0;
  double lost = 0.0;
  int* claim = (int*)calloc(area, sizeof(int));
  const int n = cycles;
  swo_particle* q = (swo_particle*)malloc(sizeof(swo_particle) * (size_t)(n + 1));
  char* alive = (char*)malloc((size_t)n + 1);
  int* home = (int*)malloc(sizeof(int) * (size_t)(n + 1));
  for (int i = 0; i < n; i++) { swo_spawn(w, &q[i]); alive[i] = 1; }
  int live = n, stamp = 0;
  /* T = tile edge, or TX * 1000 + TZ for a rectangular tile (TX rows along x, the wind axis) */
  const int TX = T >= 1000 ? T / 1000 : T, TZ = T >= 1000 ? T % 1000 : T;
  const int ntz = nz / TZ + 2;
  while (live > 0) {
    const int offx = (int)(sr % 3) * (TX / 3), offz = (int)(sr % 3) * (TZ / 3);   /* three tilings: a cell lies on the x-ring of at most one and on the z-ring of at most one */
    for (int i = 0; i < n; i++) {
      home[i] = -1;
      if (!alive[i]) continue;
      const int ix = (int)roundf(q[i].pos[0]), iz = (int)roundf(q[i].pos[2]);
      if (oob(w, ix, iz)) { alive[i] = 0; live--; lost += (double)q[i].sediment; continue; }
      home[i] = ((ix + offx) / TX) * ntz + (iz + offz) / TZ;
    }
    for (int r = 0; r < K; r++) {
      const long long round = sr * K + r;
      for (int c = 0; c < 9; c++) {
        stamp++;
        for (int i = 0; i < n; i++) {
          if (!alive[i] || home[i] < 0) continue;
          if (stagger > 0) {
            const long long start = ((long long)i * stagger) / n;
            if ((flags & 1) ? start > round : start >= (sr + 1) * K) continue;
          }
          const int ix = (int)roundf(q[i].pos[0]), iz = (int)roundf(q[i].pos[2]);
          if (oob(w, ix, iz)) { alive[i] = 0; live--; lost += (double)q[i].sediment; continue; }
          if ((ix % 3) * 3 + iz % 3 != c) continue;
          if (((ix + offx) / TX) * ntz + (iz + offz) / TZ != home[i]) continue;            /* left its tile: parked */
          const int lx = (ix + offx) % TX, lz = (iz + offz) % TZ;
          if ((lx == 0 && ix != 0) || (lx == TX - 1 && ix != nx - 1) || (lz == 0 && iz != 0) || (lz == TZ - 1 && iz != nz - 1))
            continue;                                                                 /* border ring: parked */
          const size_t ind = (size_t)ix * nz + iz;
          if (claim[ind] == stamp) continue;
          claim[ind] = stamp;
          int cell;
          if (!fly_nc(w, &q[i], NULL, &cell)) { alive[i] = 0; live--; lost += (double)q[i].sediment; continue; }
          steps++;
          swo_cascade(w, ix, iz);
        }
      }
    }
    sr++;
    if (getenv("EMUL_DBG")) fprintf(stderr, "sr %lld live %d steps %lld\n", sr, live, steps);
  }
  const float lr = w->p.lrate;
  for (int kk = 1; kk < 5; kk++) {
    float* f = w->plane[kk]; const float* t = w->plane[kk + 4];
    for (size_t i = 0; i < area; i++) f[i] = (1.0f - lr) * f[i] + lr * t[i];
  }
  free(claim); free(q); free(alive); free(home);
  if (lost_out) *lost_out = lost;
  if (rounds_out) *rounds_out = sr * K;
  return steps;
}

/* Option E: tile-resident super-rounds with BLOCK CLAIMS instead of colour phases (csrc/swe_tile.cu, claim mode).
 * An ATTEMPT serves particles of all colours at once: every serviceable particle (interior of its home tile, as in
 * option D) bids for the 3x3-lattice cells its 3x3 block touches (1, 2 or 4 of them) with a priority that is unique
 * per particle and reshuffled every attempt; a particle whose bid is the highest on ALL its lattice cells flies one
 * step + cascade, the others wait for the next attempt. Winners have disjoint blocks, so an attempt is conflict-free
 * and equivalent to processing its winners one after another. Particle i becomes active at attempt i * stagger / n
 * (granularity: the super-round of A attempts that contains it). */
static unsigned long long emul_prio(unsigned i, unsigned t) {
  unsigned long long z = ((unsigned long long)i << 32 | t) + 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
  return (z << 24) | i;   /* unique per particle */
}
long long emul_erode_claims(swo_world* w, int cycles, int stagger, int T, int A, int flags, double* lost_out,
                            long long* rounds_out) {
  const int nx = w->nx, nz = w->nz;
  const size_t area = (size_t)nx * nz;
  for (int k = 5; k < 9; k++) memset(w->plane[k], 0, sizeof(float) * area);
  long long steps = 0, sr = 0;
  double lost = 0.0;
  const int L = (flags & 2) ? 2 : 3;                      /* claim lattice cell edge */
  const int lnx = nx / L + 2, lnz = nz / L + 2;
  unsigned long long* claim = (unsigned long long*)calloc((size_t)lnx * lnz, sizeof(unsigned long long));
  const int n = cycles;
  swo_particle* q = (swo_particle*)malloc(sizeof(swo_particle) * (size_t)(n + 1));
  char* alive = (char*)malloc((size_t)n + 1);
  int* home = (int*)malloc(sizeof(int) * (size_t)(n + 1));
  int* cand = (int*)malloc(sizeof(int) * (size_t)(n + 1));
  for (int i = 0; i < n; i++) { swo_spawn(w, &q[i]); alive[i] = 1; }
  int live = n;
  const int ntz = nz / T + 2;
  while (live > 0) {
    const int off = (int)(sr % 3) * (T / 3);
    for (int i = 0; i < n; i++) {
      home[i] = -1;
      if (!alive[i]) continue;
      const int ix = (int)roundf(q[i].pos[0]), iz = (int)roundf(q[i].pos[2]);
      if (oob(w, ix, iz)) { alive[i] = 0; live--; lost += (double)q[i].sediment; continue; }
      if (stagger > 0 && ((long long)i * stagger) / n >= (sr + 1) * A) continue;
      home[i] = ((ix + off) / T) * ntz + (iz + off) / T;
    }
    for (int a = 0; a < A; a++) {
      const unsigned t = (unsigned)(sr * A + a);
      int nc = 0;
      for (int i = 0; i < n; i++) {
        if (!alive[i] || home[i] < 0) continue;
        const int ix = (int)roundf(q[i].pos[0]), iz = (int)roundf(q[i].pos[2]);
        if (oob(w, ix, iz)) { alive[i] = 0; live--; lost += (double)q[i].sediment; continue; }
        if (((ix + off) / T) * ntz + (iz + off) / T != home[i]) continue;
        const int lx = (ix + off) % T, lz = (iz + off) % T;
        if ((lx == 0 && ix != 0) || (lx == T - 1 && ix != nx - 1) || (lz == 0 && iz != 0) || (lz == T - 1 && iz != nz - 1)) continue;
        cand[nc++] = i;
        const unsigned long long pr = ((unsigned long long)(t + 1) << 40) ^ 0 , key = emul_prio((unsigned)i, t);
        (void)pr;
        for (int cx = (ix - 1 < 0 ? 0 : ix - 1) / L; cx <= (ix + 1) / L; cx++)
          for (int cz = (iz - 1 < 0 ? 0 : iz - 1) / L; cz <= (iz + 1) / L; cz++) {
            unsigned long long* c = &claim[(size_t)cx * lnz + cz];
            if (*c < key) *c = key;
          }
      }
      for (int j = 0; j < nc; j++) {
        const int i = cand[j];
        const int ix = (int)roundf(q[i].pos[0]), iz = (int)roundf(q[i].pos[2]);
        const unsigned long long key = emul_prio((unsigned)i, t);
        int won = 1;
        for (int cx = (ix - 1 < 0 ? 0 : ix - 1) / L; cx <= (ix + 1) / L; cx++)
          for (int cz = (iz - 1 < 0 ? 0 : iz - 1) / L; cz <= (iz + 1) / L; cz++)
            if (claim[(size_t)cx * lnz + cz] != key) won = 0;
        cand[j] = won ? i : -1;
      }
      /* clear the bids, then the winners step (disjoint blocks: order irrelevant) */
      memset(claim, 0, (size_t)lnx * lnz * sizeof(unsigned long long));
      for (int j = 0; j < nc; j++) {
        const int i = cand[j];
        if (i < 0) continue;
        const int ix = (int)roundf(q[i].pos[0]), iz = (int)roundf(q[i].pos[2]);
        int cell;
        if (!fly_nc(w, &q[i], NULL, &cell)) { alive[i] = 0; live--; lost += (double)q[i].sediment; continue; }
        steps++;
        swo_cascade(w, ix, iz);
      }
    }
    sr++;
    if (getenv("EMUL_DBG") && sr % 8 == 0) fprintf(stderr, "sr %lld live %d steps %lld\n", sr, live, steps);
  }
  const float lr = w->p.lrate;
  for (int kk = 1; kk < 5; kk++) {
    float* f = w->plane[kk]; const float* t = w->plane[kk + 4];
    for (size_t i = 0; i < area; i++) f[i] = (1.0f - lr) * f[i] + lr * t[i];
  }
  free(claim); free(q); free(alive); free(home); free(cand);
  if (lost_out) *lost_out = lost;
  if (rounds_out) *rounds_out = sr * A;
  return steps;
}
