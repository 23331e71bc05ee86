// Shared helpers of libnrrt.so (sm_100a only).  Error plumbing, the Philox sample stream, exact-rounding fp helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nrrt.h"

// ---- error plumbing (thread-local message, never throws across the ABI) ------------------------------------------
int nrrt_fail(int code, const char* fmt, ...);

#define NRRT_CUDA_TRY(expr)                                                                         \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return nrrt_fail(NRRT_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                             __FILE__, __LINE__);                                                   \
    } while (0)

#define NRRT_REQUIRE(cond, ...)                                  \
    do {                                                         \
        if (!(cond)) return nrrt_fail(NRRT_ERR_BAD_ARG, __VA_ARGS__); \
    } while (0)

// ---- Philox-4x32-10 sample stream: same definition as oracle/philox_stream.py -----------------------
__device__ __forceinline__ void nrrt_philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// draw k of stream `task` under `seed`: 53-bit uniform in [0,1)
__device__ __forceinline__ double nrrt_uniform(uint64_t seed, uint32_t task, uint64_t k) {
    const uint64_t blk = k >> 1;
    uint32_t c[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), task, 0x4E525254u};
    nrrt_philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t a = (k & 1) ? c[2] : c[0], b = (k & 1) ? c[3] : c[1];
    const uint64_t r53 = ((uint64_t)(a >> 5) << 26) + (uint64_t)(b >> 6);
    return __ddiv_rn((double)r53, 9007199254740992.0);
}

// ---- exact-rounding helpers: every product and sum rounded separately, never contracted ---------------
template <int DIM>
__device__ __forceinline__ double nrrt_sqdist(const double (&u)[DIM], const double (&v)[DIM]) {
    const double d0 = __dsub_rn(u[0], v[0]), d1 = __dsub_rn(u[1], v[1]);
    double s = __dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1));
    if (DIM == 3) {
        const double d2 = __dsub_rn(u[DIM - 1], v[DIM - 1]);
        s = __dadd_rn(s, __dmul_rn(d2, d2));
    }
    return s;
}

#define NRRT_K_LO 0x1.ffffffffffff8p-1 /* 1 - 2^-50 */
#define NRRT_K_HI 0x1.0000000000004p+0 /* 1 + 2^-50 */
