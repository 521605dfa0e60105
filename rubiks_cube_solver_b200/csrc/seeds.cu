// K0 -- identically seeded scrambles on the device (sm_100a): the move indices the reference draws in
//   reset(seed, k):  np.random.seed(seed); np.random.randint(A, size=k)          cube_env.py:62-65
// i.e. np.random.RandomState(seed).randint(A, size=k), for n seeds at once, so that seeded batches
// (config 1: seeds 0..1023; validation: seed = 10 * cube, train.py:180) need no host RNG and no
// upload of move bytes.
//
// NumPy's legacy generator is MT19937 seeded by init_genrand (mt[0] = seed, mt[j] = 1812433253 *
// (mt[j-1] ^ mt[j-1] >> 30) + j) and randint(A) draws 32-bit outputs, masks them with the smallest
// 2^b - 1 >= A - 1 and rejects values > A - 1 (numpy/random/_bounded_integers: masked rejection).
// Output i < 227 of a freshly seeded generator only needs mt[i], mt[i+1] and mt[i+397] of the SEED
// state, and those follow from a one-word recurrence -- so a thread keeps two running words (one at i,
// one 397 ahead) and never materialises the 624-word state.  227 raw draws cover depth <= 128 with
// > 7 sigma to spare (acceptance 3/4); a row that would need more is counted in counters[3] and padded
// with CUBE_NOOP (the host wrapper raises).
#include <cuda_runtime.h>
#include "cube_kernels.h"
#include "../../include/cube_b200.h"

namespace {

__device__ __forceinline__ uint32_t mt_next_seed_word(uint32_t prev, uint32_t j)
{
    return 1812433253u * (prev ^ (prev >> 30)) + j;
}

__global__ void __launch_bounds__(128)
seeded_moves_kernel(const uint32_t* __restrict__ seeds, long long n, int depth, uint32_t max_value, uint32_t mask,
                    uint8_t* __restrict__ moves, unsigned long long* __restrict__ counters)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t seed = seeds[i];
    uint32_t z = seed;
    for (uint32_t j = 1; j <= 397; ++j) z = mt_next_seed_word(z, j);            // mt[397]
    uint32_t x = seed, xn = mt_next_seed_word(seed, 1), idx = 0;                // mt[0], mt[1]
    uint8_t* row = moves + i * depth;
    int produced = 0;
    while (produced < depth && idx < 227) {
        const uint32_t y = (x & 0x80000000u) | (xn & 0x7fffffffu);
        uint32_t v = z ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        v ^= v >> 11;
        v ^= (v << 7) & 0x9d2c5680u;
        v ^= (v << 15) & 0xefc60000u;
        v ^= v >> 18;
        ++idx;
        x = xn;
        xn = mt_next_seed_word(xn, idx + 1);
        z = mt_next_seed_word(z, idx + 397);                                      // only used while idx + 397 <= 623
        const uint32_t val = v & mask;
        if (val <= max_value) row[produced++] = (uint8_t)val;
    }
    if (produced < depth) {
        for (int k = produced; k < depth; ++k) row[k] = CUBE_NOOP;
        if (counters) atomicAdd(&counters[3], 1ull);
    }
}

}  // namespace

namespace cube {

int launch_seeded_moves(int size, const uint32_t* seeds, long long n, int depth, uint8_t* moves,
                        unsigned long long* counters, cudaStream_t stream)
{
    if (n == 0 || depth == 0) return 0;
    const uint32_t max_value = size == 3 ? 11u : 5u;
    uint32_t mask = max_value;                                                   // smallest 2^b - 1 >= max_value
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    const long long blocks = (n + 127) / 128;
    seeded_moves_kernel<<<(unsigned)blocks, 128, 0, stream>>>(seeds, n, depth, max_value, mask, moves, counters);
    return (int)cudaGetLastError();
}

}  // namespace cube
