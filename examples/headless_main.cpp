// Headless equivalent of the reference's main (SimpleWindErosion.cpp:14-42,211-231) against the
// drop-in façade include/swe_world.hpp: seed, reserve + generate the map, then erode cycles --
// TinyEngine/OpenGL rendering is an optional consumer and absent here.
//
//   g++ -std=c++17 -O2 -Iinclude examples/headless_main.cpp -Lsimplewinderosion_b200 -lswe_b200
//       -Wl,-rpath,$PWD/simplewinderosion_b200 -o headless
//   ./headless <size> <seed> <frames> <cycles> <mode: 0 deterministic | 1 parallel> [noise layers]
#include <cstdio>
#include <cstdlib>

#include "swe_world.hpp"

int main(int argc, char* args[]) {
  const int size = argc > 1 ? std::atoi(args[1]) : 512;           // quad::tilesize (cellpool.h:186)
  const unsigned seed = argc > 2 ? (unsigned)std::atoi(args[2]) : 1;
  const int frames = argc > 3 ? std::atoi(args[3]) : 5;
  const int cycles = argc > 4 ? std::atoi(args[4]) : 8 * size;    // world.erode(8*quad::tilesize), :217
  World::mode = argc > 5 ? std::atoi(args[5]) : SWE_MODE_PARALLEL;
  const int noise_layers = argc > 6 ? std::atoi(args[6]) : 0;     // cellpool.h:388-411 (commented out there)
  try {
    World::SEED = seed;
    World::init(size, size);          // cellpool.reserve(quad::area)            :36
    World::seed(seed);                // srand(SEED)                             :27-34
    World::map.init(World::SEED, noise_layers);   // World::map.init(vertexpool, cellpool, SEED) :38
    for (int n = 0; n < frames; n++) {
      World::erode(cycles);           // :217
      World::map.sync_render_fields_from_device();   // what updatenode / windmap read, :219-230
    }
    swe_stats st;
    uint64_t hash = 0;
    swe::check(swe_get_stats(World::map.engine, &st));
    swe::check(swe_hash_height(World::map.engine, &hash));
    const quad::cell* c = World::map.getCell(swe::ivec2(size / 4, size / 2));
    std::printf("steps=%llu oob=%llu age=%llu hash=%016llx centre_height=%.9g centre_momentumx=%.9g\n",
                (unsigned long long)st.steps, (unsigned long long)st.exits_oob,
                (unsigned long long)st.exits_age, (unsigned long long)hash, c->height, c->momentumx);
    World::release();
  } catch (const swe::Error& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
