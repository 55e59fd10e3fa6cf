// TEST HELPER (host): exposes csrc/swe_block.cuh's register cascade to ctypes so that the CPU suite can
// check it against the oracle without a GPU. Build: g++ -O2 -ffp-contract=off -shared -fPIC.
#include <cmath>
#include "../../simplewinderosion_b200/csrc/swe_block.cuh"

extern "C" unsigned block_cascade(float* c, float* h8, unsigned valid, float ortho, float diag, float settling,
                                  int layered, float rock_c, const float* rk8) {
  float h[8], v[8], rk[8];
  int s[8];
  for (int k = 0; k < 8; k++) { h[k] = h8[k]; rk[k] = rk8 ? rk8[k] : 0.0f; }
  swe::block::Thr T{ortho, diag, settling};
  unsigned moved = layered ? swe::block::cascade_sorted<true>(*c, h, valid, T, v, s, rock_c, rk)
                           : swe::block::cascade_sorted<false>(*c, h, valid, T, v, s, rock_c, rk);
  unsigned changed = 0;
  for (int i = 0; i < 8; i++)
    if ((moved >> i) & 1u) { h8[s[i]] = v[i]; changed |= 1u << s[i]; }
  return changed | (moved ? 256u : 0u);
}
