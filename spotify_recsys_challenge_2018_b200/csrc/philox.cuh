// Philox4x32-10 (Salmon et al., SC'11): device twin of oracle/philox.py.
// Every dropout decision is keyed by (seed, stream, step, row, col) so masks are reproducible,
// independent of the batch sharding, and re-creatable by the CPU oracle.
#pragma once
#include <stdint.h>

namespace dae {

constexpr uint32_t kStreamInput = 0;   // input dropout   (DAEs.py:40)
constexpr uint32_t kStreamHidden = 1;  // hidden dropout  (DAEs.py:68)
constexpr uint32_t kStreamTitle = 2;   // title features  (Char_CNN.py:67)
constexpr uint32_t kStreamInit = 16;   // xavier init streams 16.. (DAEs.py:54-55)

__device__ __forceinline__ uint32_t philox_word0(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                 uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c0;
}

// u in [0,1) with 24 random bits
__device__ __forceinline__ float philox_uniform24(unsigned long long seed, uint32_t stream, unsigned long long step,
                                                  uint32_t row, uint32_t col) {
    const uint32_t c2 = static_cast<uint32_t>(step);
    const uint32_t c3 = (static_cast<uint32_t>((step >> 32) & 0xFFFFFFull) << 8) | (stream & 0xFFu);
    const uint32_t w = philox_word0(col, row, c2, c3, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    return static_cast<float>(w >> 8) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ bool philox_keep(unsigned long long seed, uint32_t stream, unsigned long long step,
                                            uint32_t row, uint32_t col, float keep_prob) {
    return philox_uniform24(seed, stream, step, row, col) < keep_prob;
}

}  // namespace dae
