/* sweep_stats.c -- design probe (not product, not oracle): how much work does a whole-map cascade
 * sweep really have near the repose angle?  Runs nine-colour sweeps (world.hpp:79-153 semantics,
 * IEEE fp32) on a U(0,1) map and prints, per sweep:
 *   moved   fraction of centres whose World::cascade changes a value
 *   hot     fraction of cells that are an endpoint of a pair with excess >= ulp(min|h|)  (the
 *           conservative "may move" pair test, each pair visited once)
 *   cand    fraction of centres that are hot or have a changed cell in their 3x3 block
 *   coarse  fraction of centres with some pair excess > 0
 * build: gcc -O2 -ffp-contract=off -o build/sweep_stats tools/sweep_stats.c -lm */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int N;
static float *H;
static unsigned char *hot, *chg;
static const int DX[8] = {-1, -1, -1, 0, 0, 1, 1, 1}, DZ[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
static float TO, TD;

static float ulp_of(float a) {
  uint32_t u; a = fabsf(a); memcpy(&u, &a, 4);
  int e = (int)(u >> 23) - 23;
  if (e <= 0) return 0.0f;
  u = (uint32_t)e << 23; memcpy(&a, &u, 4); return a;
}
static float MARGIN = 1.0f;
static int pair_hot(float a, float b, float t) {
  float e = fabsf(a - b) - t;
  float u = fminf(ulp_of(a), ulp_of(b)) * MARGIN;
  return e > 0.0f && e >= u;
}

static int cascade(int x, int z) {
  float h[8]; int id[8], k[8], num = 0;
  for (int i = 0; i < 8; i++) {
    int a = x + DX[i], b = z + DZ[i];
    if (a < 0 || b < 0 || a >= N || b >= N) continue;
    h[num] = H[a * N + b]; id[num] = a * N + b; k[num] = i; num++;
  }
  int ord[8];
  for (int i = 0; i < num; i++) ord[i] = i;
  for (int i = 1; i < num; i++) { int v = ord[i], j = i; while (j > 0 && h[v] > h[ord[j - 1]]) { ord[j] = ord[j - 1]; j--; } ord[j] = v; }
  int moved = 0;
  for (int i = 0; i < num; i++) {
    int q = ord[i];
    float c = H[x * N + z];
    float diff = c - h[q];
    if (diff == 0) continue;
    float t = (DX[k[q]] && DZ[k[q]]) ? TD : TO;
    float excess = fabsf(diff) - t;
    if (excess <= 0) continue;
    float tr = 1.0f * excess / 2.0f;
    float nc, nh;
    if (diff > 0) { nc = c - tr; nh = H[id[q]] + tr; } else { nc = c + tr; nh = H[id[q]] - tr; }
    if (nc != c) { chg[x * N + z] = 1; moved = 1; }
    if (nh != H[id[q]]) { chg[id[q]] = 1; moved = 1; }
    H[x * N + z] = nc; H[id[q]] = nh;
  }
  return moved;
}

int main(int argc, char** argv) {
  N = argc > 1 ? atoi(argv[1]) : 512;
  int sweeps = argc > 2 ? atoi(argv[2]) : 200;
  if (argc > 3) MARGIN = (float)atof(argv[3]);
  H = malloc(sizeof(float) * N * N); hot = malloc(N * N); chg = malloc(N * N);
  uint64_t s = 42;
  for (int i = 0; i < N * N; i++) { s = s * 6364136223846793005ull + 1442695040888963407ull; H[i] = (float)((s >> 40) & 0xffffff) / 16777216.0f; }
  TO = 0.005f; TD = sqrtf(2.0f) * 0.005f;
  for (int sw = 0; sw < sweeps; sw++) {
    memset(hot, 0, N * N); memset(chg, 0, N * N);
    long ncoarse = 0;
    for (int x = 0; x < N; x++) for (int z = 0; z < N; z++) {
      float a = H[x * N + z]; int co = 0;
      for (int i = 0; i < 8; i++) {
        int p = x + DX[i], q = z + DZ[i];
        if (p < 0 || q < 0 || p >= N || q >= N) continue;
        float t = (DX[i] && DZ[i]) ? TD : TO;
        if (pair_hot(a, H[p * N + q], t)) hot[x * N + z] = 1;
        if (fabsf(a - H[p * N + q]) - t > 0) co = 1;
      }
      ncoarse += co;
    }
    long nhot = 0, ncand = 0, nmoved = 0, miss = 0;
    for (int i = 0; i < N * N; i++) nhot += hot[i];
    for (int c = 0; c < 9; c++) {
      int cx = c / 3, cz = c % 3;
      for (int x = cx; x < N; x += 3) for (int z = cz; z < N; z += 3) {
        int cand = hot[x * N + z];
        for (int a = x - 1; a <= x + 1 && !cand; a++) for (int b = z - 1; b <= z + 1; b++)
          if (a >= 0 && b >= 0 && a < N && b < N && chg[a * N + b]) { cand = 1; break; }
        ncand += cand;
        int mv = cascade(x, z);
        nmoved += mv;
        if (mv && !cand) miss++;
      }
    }
    float mx = -1;
    for (int x = 0; x < N; x++) for (int z = 0; z + 1 < N; z++) { float e = fabsf(H[x * N + z] - H[x * N + z + 1]) - TO; if (e > mx) mx = e; }
    for (int x = 0; x + 1 < N; x++) for (int z = 0; z < N; z++) {
      float e = fabsf(H[x * N + z] - H[(x + 1) * N + z]) - TO; if (e > mx) mx = e;
      if (z > 0) { e = fabsf(H[x * N + z] - H[(x + 1) * N + z - 1]) - TD; if (e > mx) mx = e; }
      if (z + 1 < N) { e = fabsf(H[x * N + z] - H[(x + 1) * N + z + 1]) - TD; if (e > mx) mx = e; }
    }
    double nn = (double)N * N;
    if (sw < 5 || sw % 10 == 0 || sw == sweeps - 1)
      printf("sweep %3d  moved %.4f  hot %.4f  cand %.4f  coarse %.4f  miss %ld  maxexcess %.3g\n", sw, nmoved / nn,
             nhot / nn, ncand / nn, ncoarse / nn, miss, mx);
  }
  return 0;
}
