// wgrad_tc.cu — weight gradient of the 3x3 / 1x1 stride-1 convolutions on the sm_100a tensor cores (the
// backward pass behind UNet_cooler.training_step, train.py:205-210 -> loss.backward()).
//
//   dW[co][ky][kx][ci] = sum over (n, y, x) of  dY[n, y, x, co] * X[n, y + ky - 1, x + kx - 1, ci]
//
// GEMM view: M = 128 output channels, N = BLOCK_N input channels, K = PIXELS.  Two operand paths share the kernel:
//   MN = true  (nrrt_conv_wgrad_nhwc, the product path): both operands are read straight from the channels-last buffers as
//     MN-major UMMA operands.  A TMA box of 64 pixels x 64 channels is 64 K-rows of 128 bytes = eight 1024-byte SWIZZLE_128B
//     atoms; both tap shifts are TMA coordinates of OUTER dimensions and out-of-range pixels are zero-filled = the padding.
//     Nothing is transposed or copied.
//   MN = false (nrrt_conv_wgrad, kept as cross-check): operands channel-major ([C][N][H][W], written by
//     nrrt_to_channel_major), K-major tiles.  The kx shift is then along the innermost dimension, and a swizzled TMA row that
//     starts at an odd pixel is an illegal instruction (measured): X is written as three copies shifted by -1 / 0 / +1 pixel.
// One CTA owns (128 co) x (BLOCK_N ci) x (one kx, all ky) and walks image-column strips of 64 pixels row by row: the
// operand row X[y + 1] loaded for k-block y is the ky = 2 operand of k-block y, the ky = 1 operand of k-block y + 1 and the
// ky = 0 operand of k-block y + 2 -- each k-block loads ONE A tile and ONE B tile for three MMAs (nky accumulators of
// BLOCK_N TMEM columns).  K is split over CTAs (image, column block, row segment); partial sums are added to the fp32
// gradient with red.global.add.v4.f32.
// Warp roles (192 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue (TMEM -> red.global.add.f32).
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "nrrt_common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int kWgThreads = 192;
constexpr int kNA = 4;   // A ring (dY rows)
constexpr int kNB = 6;   // B ring (X rows): nky live + prefetch

struct WgradParams {
    int H, xblocks, strips;       // work units: strips = N * xblocks * segs (image, 64-pixel column block, row segment)
    int seg_rows, segs;           // rows per segment, segments per column
    int co, ci;                   // true channel counts
    int nky, nkx, pad;            // 3,3,1 or 1,1,0
    int co_blocks, ci_blocks, splits;
    float* dw;                    // fp32 [co][nky * nkx][ci_stride]
    int ci_stride;
    int dbg;                      // NRRT_DEBUG_SWITCHES builds only: 1 = no MMA, 2 = no TMEM loads, 4 = no TMA
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

__device__ __forceinline__ void tma_load_5d_(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}

template <int BLOCK_N>
struct WgLayout {
    static constexpr uint32_t a_stage = 128 * 128;
    static constexpr uint32_t b_stage = BLOCK_N * 128;
    static constexpr uint32_t off_b = kNA * a_stage;
    static constexpr uint32_t off_bar = off_b + kNB * b_stage;
    static constexpr uint32_t off_tmem = off_bar + (2 * kNA + 2 * kNB + 1) * 8;
    // at least 120 KB: two CTAs must never share an SM (each allocates all 512 tensor-memory columns; the second would wait forever)
    static constexpr uint32_t need = off_tmem + 16 + 1024;
    static constexpr uint32_t total = need > 120u * 1024u ? need : 120u * 1024u;
};

// MN-major SWIZZLE_128B operand descriptor (channels-last tiles: a K row = one pixel = 128 bytes of 64 channels; 8 pixels
// form a 1024-byte swizzle atom): SBO = 1024 B between 8-pixel groups, LBO = 8192 B between 64-channel blocks (a block is
// one TMA box of 64 pixels x 64 channels).  cute::UMMA canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(8192 >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// MN = true: both operands are read straight from the channels-last activation / gradient buffers (MN-major UMMA operands:
// no layout kernel, no shifted copies -- a tap shift in either direction is a TMA coordinate of an OUTER dimension).
template <int BLOCK_N, bool MN = false>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
             const __grid_constant__ WgradParams p) {
    using L = WgLayout<BLOCK_N>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar0 = base + L::off_bar;
    auto full_a = [&](int s) { return bar0 + 8u * s; };
    auto empty_a = [&](int s) { return bar0 + 8u * (kNA + s); };
    auto full_b = [&](int s) { return bar0 + 8u * (2 * kNA + s); };
    auto empty_b = [&](int s) { return bar0 + 8u * (2 * kNA + kNB + s); };
    const uint32_t tfull = bar0 + 8u * (2 * kNA + 2 * kNB);
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + L::off_tmem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    int t = blockIdx.x;
    const int kx = t % p.nkx; t /= p.nkx;
    const int nb = t % p.ci_blocks; t /= p.ci_blocks;
    const int cb = t % p.co_blocks; t /= p.co_blocks;
    const int split = t;
    const int per = (p.strips + p.splits - 1) / p.splits;
    const int s_begin = split * per, s_end = min(p.strips, s_begin + per);

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kNA; ++s) { mbar_init(full_a(s), 1); mbar_init(empty_a(s), 1); }
        for (int s = 0; s < kNB; ++s) { mbar_init(full_b(s), 1); mbar_init(empty_b(s), 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + L::off_tmem), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t a_seq = 0, b_seq = 0;
            for (int s = s_begin; s < s_end; ++s) {
                const int seg = s % p.segs, sx = s / p.segs;
                const int n = sx / p.xblocks, x0 = (sx % p.xblocks) * 64;
                const int y0 = seg * p.seg_rows, rows = min(p.seg_rows, p.H - y0);
                for (int r = y0 - p.pad; r < y0 + rows + p.pad; ++r) {
                    {
                        const uint32_t slot = b_seq % kNB, ph = (b_seq / kNB) & 1u;
                        mbar_wait(empty_b(slot), ph ^ 1u);
                        mbar_expect_tx(full_b(slot), (p.dbg & 4) ? 0u : L::b_stage);
                        if (MN) {
#pragma unroll
                            for (int j = 0; j < BLOCK_N / 64; ++j)
                                tma_load_4d(base + L::off_b + slot * L::b_stage + j * 8192, &map_b, full_b(slot), nb * BLOCK_N + 64 * j,
                                            x0 + kx - p.pad, r, n);
                        } else if (!(p.dbg & 4))   // the kx shift lives in the operand copy (TMA cannot start a swizzled row off a 16-byte boundary)
                        tma_load_5d_(base + L::off_b + slot * L::b_stage, &map_b, full_b(slot), x0, r, n, nb * BLOCK_N, kx);
                        ++b_seq;
                    }
                    const int y = r - p.pad;
                    if (y >= y0) {
                        const uint32_t slot = a_seq % kNA, ph = (a_seq / kNA) & 1u;
                        mbar_wait(empty_a(slot), ph ^ 1u);
                        mbar_expect_tx(full_a(slot), (p.dbg & 4) ? 0u : L::a_stage);
                        if (MN) {
                            tma_load_4d(base + slot * L::a_stage, &map_a, full_a(slot), cb * 128, x0, y, n);
                            tma_load_4d(base + slot * L::a_stage + 8192, &map_a, full_a(slot), cb * 128 + 64, x0, y, n);
                        } else if (!(p.dbg & 4))
                        tma_load_4d(base + slot * L::a_stage, &map_a, full_a(slot), x0, y, n, cb * 128);
                        ++a_seq;
                    }
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc(BLOCK_N) | (MN ? ((1u << 15) | (1u << 16)) : 0u);   // bits 15 / 16: A / B are MN-major
        constexpr uint64_t kstep = MN ? (2048 >> 4) : 2;   // one UMMA_K = 16 pixels: two 1024-byte atoms (MN-major) / 32 bytes (K-major)
        uint32_t a_seq = 0, bq = 0, b_waited = 0;   // b_waited = number of B loads already waited for (they complete in order)
        bool first = true;
        for (int s = s_begin; s < s_end; ++s) {
            const int rows = min(p.seg_rows, p.H - (s % p.segs) * p.seg_rows);
            for (int y = 0; y < rows; ++y) {     // y = row inside the segment
                const uint32_t a_slot = a_seq % kNA;
                mbar_wait(full_a(a_slot), (a_seq / kNA) & 1u);
                const uint32_t need = bq + (uint32_t)(y + 2 * p.pad);   // last B load this k-block reads
                while (b_waited <= need) {
                    mbar_wait(full_b(b_waited % kNB), (b_waited / kNB) & 1u);
                    ++b_waited;
                }
                tc_fence_after();
                const uint64_t adesc = MN ? make_smem_desc_mn(base + a_slot * L::a_stage) : make_smem_desc(base + a_slot * L::a_stage);
                if (elect_one()) {
                    for (int ky = 0; ky < p.nky && !(p.dbg & 1); ++ky) {
                        const uint32_t b_slot = (bq + (uint32_t)(y + ky)) % kNB;
                        const uint64_t bdesc = MN ? make_smem_desc_mn(base + L::off_b + b_slot * L::b_stage)
                                                  : make_smem_desc(base + L::off_b + b_slot * L::b_stage);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            tc_mma_f16(tmem_base + (uint32_t)(ky * BLOCK_N), adesc + kstep * kk, bdesc + kstep * kk, idesc,
                                       (first && kk == 0) ? 0u : 1u);
                    }
                    tc_commit(empty_a(a_slot));
                    tc_commit(empty_b((bq + (uint32_t)y) % kNB));          // operand row y - pad is not needed again
                    if (y == rows - 1)
                        for (int e = 1; e <= 2 * p.pad; ++e) tc_commit(empty_b((bq + (uint32_t)(y + e)) % kNB));
                }
                __syncwarp();
                first = false;
                ++a_seq;
            }
            bq += (uint32_t)(rows + 2 * p.pad);
        }
        if (elect_one()) tc_commit(tfull);
        __syncwarp();
    } else {
        const int q = warp & 3;                    // TMEM lane quarter this warp may read (warp id % 4)
        const int co = cb * 128 + q * 32 + lane;
        if (s_begin < s_end) {
            mbar_wait(tfull, 0u);
            tc_fence_after();
            const int taps = p.nky * p.nkx;
            for (int ky = 0; ky < p.nky; ++ky) {
                float* row = p.dw + ((size_t)co * taps + (ky * p.nkx + kx)) * p.ci_stride + nb * BLOCK_N;
#pragma unroll 1
                for (int c = 0; c < BLOCK_N / 32; ++c) {
                    uint32_t v[32];
                    if (!(p.dbg & 2))
                    tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ky * BLOCK_N + c * 32), v);
                    if (co < p.co) {
                        const int ci0 = nb * BLOCK_N + c * 32;
                        if (ci0 + 32 <= p.ci && (p.ci_stride & 3) == 0) {   // 8 vector reductions of 16 bytes instead of 32 scalar ones
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + c * 32 + j), "f"(__uint_as_float(v[j])),
                                             "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                                             : "memory");
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (ci0 + j < p.ci) atomicAdd(row + c * 32 + j, __uint_as_float(v[j]));
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn wg_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

// channel-major fp16 tensor [copies][C][N][H][W] -> map (W, H, N, C[, copies]), box (64, 1, 1, rows[, 1]), 128-byte swizzle, zero fill
int make_cm_map(CUtensorMap* m, const void* ptr, int C, int N, int H, int W, int rows, int copies) {
    EncodeTiledFn enc = wg_encode_fn();
    if (!enc) return nrrt_fail(NRRT_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
    const cuuint64_t dims[5] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)C, (cuuint64_t)copies};
    const cuuint64_t strides[4] = {(cuuint64_t)W * 2, (cuuint64_t)W * H * 2, (cuuint64_t)W * H * N * 2, (cuuint64_t)W * H * N * C * 2};
    const cuuint32_t box[5] = {64, 1, 1, (cuuint32_t)rows, 1};
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, copies > 0 ? 5 : 4, const_cast<void*>(ptr), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return nrrt_fail(NRRT_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for [%d][%d][%d][%d]", (int)r, C, N, H, W);
    return NRRT_OK;
}

template <int BLOCK_N, bool MN = false>
int launch_wgrad(const CUtensorMap& ma, const CUtensorMap& mb, const WgradParams& p, int grid, cudaStream_t st) {
    using L = WgLayout<BLOCK_N>;
    auto kern = wgrad_kernel<BLOCK_N, MN>;
    static int done_dev = -1;
    int dev = 0;
    NRRT_CUDA_TRY(cudaGetDevice(&dev));
    if (done_dev != dev) {
        NRRT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total));
        done_dev = dev;
    }
    kern<<<grid, kWgThreads, L::total, st>>>(ma, mb, p);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

// ---- layout kernel: channels-last slice -> channel-major, with the three pixel mappings the backward pass needs ----------
struct CmArgs {
    const __half* in;       // [N][Hin][Win][in_cstride], channels [0, C) of the slice at `in`
    const __half* mask;     // optional ReLU mask source, same grid as in: value kept where mask > 0
    __half* out;            // mode 0: [copies][C][N][H][W], copy k = the image shifted so that out[x] = in[x + k - copies / 2];
                            // 1: [C][N][2H][2W] (pre-zeroed, value at even positions);
                            // 2: [4][C][N][H/2][W/2] (group g = (y & 1) * 2 + (x & 1))
    long long in_cstride, mask_cstride;
    int N, H, W, C, mode, copies;
};

__global__ void __launch_bounds__(256) to_channel_major_kernel(const CmArgs a) {
    // tile_t[c][1 + px]: 64 channels x pixels x0 - 1 .. x0 + 64 of one image row, channel-major (33 words per row: the 8
    // channel chunks of a pixel land 8 banks apart)
    __shared__ __half tile_t[64][66];
    const int x0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
    const int n = blockIdx.z / a.H, y = blockIdx.z % a.H;
    const int tid = threadIdx.x;
    for (int i = tid; i < 66 * 8; i += 256) {          // 66 pixels x 8 chunks of 8 channels: 128 contiguous bytes per pixel
        const int px = i >> 3, ch = (i & 7) * 8, x = x0 + px - 1;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (x >= 0 && x < a.W && c0 + ch < a.C) {
            const size_t pix = ((size_t)n * a.H + y) * a.W + x;
            v = *reinterpret_cast<const uint4*>(a.in + pix * a.in_cstride + c0 + ch);
            if (a.mask) {
                const uint4 m = *reinterpret_cast<const uint4*>(a.mask + pix * a.mask_cstride + c0 + ch);
                const __half2* mh = reinterpret_cast<const __half2*>(&m);
                __half2* vh = reinterpret_cast<__half2*>(&v);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 mf = __half22float2(mh[j]);
                    float2 vf = __half22float2(vh[j]);
                    if (!(mf.x > 0.f)) vf.x = 0.f;
                    if (!(mf.y > 0.f)) vf.y = 0.f;
                    vh[j] = __floats2half2_rn(vf.x, vf.y);
                }
            }
        }
        const __half* vh = reinterpret_cast<const __half*>(&v);
#pragma unroll
        for (int j = 0; j < 8; ++j) tile_t[ch + j][px] = vh[j];
    }
    __syncthreads();
    // a thread writes 8 consecutive pixels of one channel (16 bytes); a warp writes 4 channels x 128 contiguous bytes
    for (int i = tid; i < 64 * 8; i += 256) {
        const int c = i >> 3, gp = (i & 7) * 8, x = x0 + gp;
        if (c0 + c >= a.C || x >= a.W) continue;
        const __half* src = &tile_t[c][gp + 1];
        const bool full = x + 8 <= a.W && (a.W & 7) == 0;
        if (a.mode == 0) {
            for (int k = 0; k < a.copies; ++k) {
                const int sh = k - a.copies / 2;
                __align__(16) __half v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = src[j + sh];
                __half* dst = a.out + ((((size_t)k * a.C + c0 + c) * a.N + n) * a.H + y) * a.W + x;
                if (full) *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(v);
                else for (int j = 0; j < 8 && x + j < a.W; ++j) dst[j] = v[j];
            }
        } else if (a.mode == 1) {   // row 2y: values at even x, zeros between; row 2y + 1: zeros
            __half* r0 = a.out + (((size_t)(c0 + c) * a.N + n) * (2 * a.H) + 2 * y) * (2 * a.W) + 2 * x;
            __half* r1 = r0 + 2 * a.W;
            const __half z = __float2half(0.f);
            if (full) {
                __align__(16) __half v[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) { v[2 * j] = src[j]; v[2 * j + 1] = z; }
                reinterpret_cast<uint4*>(r0)[0] = reinterpret_cast<const uint4*>(v)[0];
                reinterpret_cast<uint4*>(r0)[1] = reinterpret_cast<const uint4*>(v)[1];
                reinterpret_cast<uint4*>(r1)[0] = make_uint4(0, 0, 0, 0);
                reinterpret_cast<uint4*>(r1)[1] = make_uint4(0, 0, 0, 0);
            } else {
                for (int j = 0; j < 8 && x + j < a.W; ++j) { r0[2 * j] = src[j]; r0[2 * j + 1] = z; r1[2 * j] = z; r1[2 * j + 1] = z; }
            }
        } else {                    // space to depth: even / odd pixels of this row go to groups (y & 1) * 2 + {0, 1}
            const size_t plane = (size_t)a.C * a.N * (a.H / 2) * (a.W / 2);
            __half* d0 = a.out + (size_t)((y & 1) * 2) * plane + (((size_t)(c0 + c) * a.N + n) * (a.H / 2) + (y >> 1)) * (a.W / 2) + (x >> 1);
            __half* d1 = d0 + plane;
            if (full) {
                __align__(8) __half e[4];
                __align__(8) __half o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { e[j] = src[2 * j]; o[j] = src[2 * j + 1]; }
                *reinterpret_cast<uint2*>(d0) = *reinterpret_cast<const uint2*>(e);
                *reinterpret_cast<uint2*>(d1) = *reinterpret_cast<const uint2*>(o);
            } else {
                for (int j = 0; j < 8 && x + j < a.W; ++j) ((j & 1) ? d1 : d0)[j >> 1] = src[j];
            }
        }
    }
}

}  // namespace

extern "C" int nrrt_to_channel_major(const void* in, int64_t in_cstride, const void* relu_mask, int64_t mask_cstride, int n, int h,
                                     int w, int c, int mode, int copies, void* out, void* stream) {
    NRRT_REQUIRE(copies == 1 || (copies == 3 && mode == 0), "copies must be 1, or 3 with mode 0");
    NRRT_REQUIRE(in && out, "null pointer");
    NRRT_REQUIRE(n >= 1 && h >= 1 && w >= 1 && c >= 8 && c % 8 == 0 && in_cstride % 8 == 0 && mask_cstride % 8 == 0, "bad shape");
    NRRT_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0 (plain), 1 (zero-stuffed x2) or 2 (space to depth)");
    NRRT_REQUIRE(mode != 2 || (h % 2 == 0 && w % 2 == 0), "space-to-depth needs even sides");
    NRRT_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)relu_mask & 15) == 0, "pointers must be 16-byte aligned");
    NRRT_REQUIRE((long long)n * h <= 65535, "n * h must be <= 65535");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    NRRT_REQUIRE(((uintptr_t)out & 15) == 0, "out must be 16-byte aligned");
    CmArgs a;
    a.in = (const __half*)in; a.mask = (const __half*)relu_mask; a.out = (__half*)out;
    a.in_cstride = in_cstride; a.mask_cstride = mask_cstride; a.N = n; a.H = h; a.W = w; a.C = c; a.mode = mode; a.copies = copies;
    to_channel_major_kernel<<<dim3((w + 63) / 64, (c + 63) / 64, n * h), 256, 0, st>>>(a);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

namespace {
// channels-last fp16 slice [N][H][W][cstride], channels [0, C) at ptr -> map (C, W, H, N), box (64 channels, 64 pixels, 1, 1)
int make_cl_map(CUtensorMap* m, const void* ptr, int C, int64_t cstride, int N, int H, int W) {
    EncodeTiledFn enc = wg_encode_fn();
    if (!enc) return nrrt_fail(NRRT_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
    const cuuint64_t cs = (cuuint64_t)cstride * 2;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[3] = {cs, cs * W, cs * W * H};
    const cuuint32_t box[4] = {64, 64, 1, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return nrrt_fail(NRRT_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for channels-last [%d][%d][%d][%d]", (int)r, N, H, W, C);
    return NRRT_OK;
}

int wgrad_plan(WgradParams& p, int cout, int cin, int n, int h, int w, int ksize, float* dw, int ci_stride, int block_n, int* grid) {
    p.H = h; p.xblocks = (w + 63) / 64;   // a last column block wider than the image reads zeros (TMA fill): correct, just idle MMA rows
    p.co = cout; p.ci = cin;
    p.nky = ksize; p.nkx = ksize; p.pad = ksize / 2;
    p.co_blocks = (cout + 127) / 128; p.ci_blocks = (cin + block_n - 1) / block_n;
    p.dw = dw; p.ci_stride = ci_stride;
    p.dbg = 0;
    int dev = 0, sms = 0;
    NRRT_CUDA_TRY(cudaGetDevice(&dev));
    NRRT_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int tiles = p.co_blocks * p.ci_blocks * p.nkx;
    int seg_rows = h;
    while (seg_rows > 16 && (long long)tiles * n * p.xblocks * ((h + seg_rows - 1) / seg_rows) < 2LL * sms) seg_rows = (seg_rows + 1) / 2;
    p.seg_rows = seg_rows; p.segs = (h + seg_rows - 1) / seg_rows;
    p.strips = n * p.xblocks * p.segs;
    int splits = (2 * sms + tiles - 1) / tiles;
    if (splits > p.strips) splits = p.strips;
    if (splits < 1) splits = 1;
    p.splits = splits;
    *grid = tiles * splits;
    return NRRT_OK;
}
}  // namespace

extern "C" int nrrt_conv_wgrad_nhwc(const void* dy, int64_t dy_cstride, int cout, const void* x, int64_t x_cstride, int cin, int n, int h,
                                    int w, int ksize, float* dw, int ci_stride, void* stream) {
    NRRT_REQUIRE(dy && x && dw, "null pointer");
    NRRT_REQUIRE(ksize == 1 || ksize == 3, "ksize must be 1 or 3");
    NRRT_REQUIRE(n >= 1 && h >= 1 && w >= 1, "bad grid");
    NRRT_REQUIRE(cout >= 1 && cin >= 1 && ci_stride >= cin, "bad channel counts");
    NRRT_REQUIRE(dy_cstride % 8 == 0 && x_cstride % 8 == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)x & 15) == 0,
                 "operands must be 16-byte aligned with channel strides that are multiples of 8");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    CUtensorMap ma, mb;
    if (int rc = make_cl_map(&ma, dy, cout, dy_cstride, n, h, w)) return rc;
    // input channels in runs of 128 (N = 128 MMAs); a remainder of 64 (192, 384 + 64, ...) is a second launch with N = 64
    const int main_ci = cin >= 128 ? (cin / 128) * 128 : 0, tail_ci = cin - main_ci;
    for (int part = 0; part < 2; ++part) {
        const int c0 = part == 0 ? 0 : main_ci, cn = part == 0 ? main_ci : tail_ci;
        if (cn == 0) continue;
        const int block_n = part == 0 ? 128 : 64;
        WgradParams p;
        int grid = 0;
        if (int rc = wgrad_plan(p, cout, cn, n, h, w, ksize, dw + c0, ci_stride, block_n, &grid)) return rc;
        if (int rc = make_cl_map(&mb, (const __half*)x + c0, cn, x_cstride, n, h, w)) return rc;
        const int rc = block_n == 128 ? launch_wgrad<128, true>(ma, mb, p, grid, st) : launch_wgrad<64, true>(ma, mb, p, grid, st);
        if (rc) return rc;
    }
    return NRRT_OK;
}

extern "C" int nrrt_conv_wgrad(const void* dy_cm, int cout, const void* x_cm, int cin, int n, int h, int w, int ksize, float* dw,
                               int ci_stride, void* stream) {
    NRRT_REQUIRE(dy_cm && x_cm && dw, "null pointer");
    NRRT_REQUIRE(ksize == 1 || ksize == 3, "ksize must be 1 or 3");
    NRRT_REQUIRE(n >= 1 && h >= 1 && w >= 64 && w % 64 == 0, "image rows must be a multiple of 64 pixels (got w = %d)", w);
    NRRT_REQUIRE(cout >= 1 && cin >= 1 && ci_stride >= cin, "bad channel counts");
    NRRT_REQUIRE(((uintptr_t)dy_cm & 15) == 0 && ((uintptr_t)x_cm & 15) == 0, "operands must be 16-byte aligned");
    const int block_n = (cin % 128 == 0 || cin > 192) ? 128 : (cin % 96 == 0 ? 96 : 64);   // 192 channels: two MMAs of N = 96
    WgradParams p;
    p.H = h; p.xblocks = w / 64;
    p.co = cout; p.ci = cin;
    p.nky = ksize; p.nkx = ksize; p.pad = ksize / 2;
    p.co_blocks = (cout + 127) / 128; p.ci_blocks = (cin + block_n - 1) / block_n;
    p.dw = dw; p.ci_stride = ci_stride;
    p.dbg = 0;
#ifdef NRRT_DEBUG_SWITCHES
    p.dbg = getenv("NRRT_WG_DBG") ? atoi(getenv("NRRT_WG_DBG")) : 0;
#endif
    int dev = 0, sms = 0;
    NRRT_CUDA_TRY(cudaGetDevice(&dev));
    NRRT_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int tiles = p.co_blocks * p.ci_blocks * p.nkx;
    // K split: columns of 64 pixels, cut into row segments (>= 16 rows: a segment re-reads 2 * pad halo rows) until there are
    // about two waves of CTAs; a CTA accumulates its work units in tensor memory and adds the result once
    int seg_rows = h;
    while (seg_rows > 16 && (long long)tiles * n * p.xblocks * ((h + seg_rows - 1) / seg_rows) < 2LL * sms) seg_rows = (seg_rows + 1) / 2;
    p.seg_rows = seg_rows; p.segs = (h + seg_rows - 1) / seg_rows;
    p.strips = n * p.xblocks * p.segs;
    int splits = (2 * sms + tiles - 1) / tiles;
    if (splits > p.strips) splits = p.strips;
    if (splits < 1) splits = 1;
    p.splits = splits;
    CUtensorMap ma, mb;
    if (int rc = make_cm_map(&ma, dy_cm, cout, n, h, w, 128, 0)) return rc;
    if (int rc = make_cm_map(&mb, x_cm, cin, n, h, w, block_n, ksize)) return rc;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int grid = tiles * splits;
    return block_n == 128 ? launch_wgrad<128>(ma, mb, p, grid, st) : (block_n == 96 ? launch_wgrad<96>(ma, mb, p, grid, st) : launch_wgrad<64>(ma, mb, p, grid, st));
}
