// Philox4x32-10 counter-based streams (RNG contract: network-slicing_b200/philox.py).
// key = base seed of the batch, counter = (draw index, stream id, slice index, global env id); one tick per variate.
// (The env id is a counter word, not a key offset: batches with adjacent seeds share no streams.)
#pragma once
#include <cstdint>

namespace rs {

enum Stream : uint32_t { STREAM_RAN = 0, STREAM_CHAN = 1, STREAM_L1RX = 2, STREAM_VBR = 3, STREAM_MTC = 4, STREAM_KBRL = 5 };

struct U4 { uint32_t x, y, z, w; };

__device__ __forceinline__ U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                            uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return U4{c0, c1, c2, c3};
}

// One (env, slice, purpose) stream; `n` is the persistent draw counter.
struct PhiloxStream {
    uint32_t k0, k1, slice, stream, n, env;
    __device__ __forceinline__ U4 raw() { return philox4x32_10(n++, stream, slice, env, k0, k1); }
    // 53-bit uniform in [0,1): ((x>>5)*2^26 + (y>>6)) / 2^53, exact in fp64
    __device__ __forceinline__ double u01() {
        const U4 r = raw();
        return ((double)(r.x >> 5) * 67108864.0 + (double)(r.y >> 6)) * (1.0 / 9007199254740992.0);
    }
    __device__ __forceinline__ double exponential(double scale) { return -log(1.0 - u01()) * scale; }
    __device__ __forceinline__ uint32_t integers(uint32_t nn) { return __umulhi(raw().x, nn); }
    __device__ __forceinline__ double normal(double mu, double sigma) {
        const double u1 = u01(), u2 = u01();
        const double z = sqrt(-2.0 * log(1.0 - u1)) * cos(6.283185307179586 * u2);
        return mu + sigma * z;
    }
};

}  // namespace rs
