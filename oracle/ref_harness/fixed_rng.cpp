// TEST INFRASTRUCTURE (oracle). LD_PRELOAD shim: makes std::random_device deterministic so the reference prover randomness (r, s) is pinned.
#include <random>
#include <cstdlib>
#include <cstdint>
static uint64_t ctr = 0;
unsigned int std::random_device::_M_getval() {
    static uint64_t seed = getenv("ZK_FIXED_SEED") ? strtoull(getenv("ZK_FIXED_SEED"), 0, 0) : 1;
    uint64_t z = (seed + (++ctr) * 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    return (unsigned int)z;
}
