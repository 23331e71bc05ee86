// conv_tc.cu — implicit-GEMM convolution on the sm_100a tensor cores (tcgen05.mma, TMEM accumulators, TMA operands).
//
// Replaces the dense contraction of UNet_cooler.forward / ThreeD_UNet_cooler.forward (train.py:171-203,
// ThreeD_train.py:169-204): every Conv2d/Conv3d 3x3(x3) stride 1/2, the 1x1(x1) stride-2 ResNet downsamples and the
// ConvTranspose k=2 s=2 layers, with the eval-mode BatchNorm + bias folded into a per-channel scale/shift, the residual
// add and the ReLU fused into the epilogue.
//
// GEMM view:  D[pixel, cout] = sum_{tap, cin} X[pixel + tap, cin] * W[cout, tap, cin]
//   M = 128 output pixels per CTA tile: a (bd x bh x bw) box of the NDHWC activation tensor.  Each tap is ONE TMA tiled
//       load of that box shifted by the tap offset (out-of-bounds rows are zero-filled by TMA = the conv padding), so
//       no im2col buffer and no padded copies exist anywhere; stride-2 layers use the tensor map's element strides.
//   N = BLOCK_N output channels (64/128/256) = one tcgen05.mma instruction wide.
//   K = 64 input channels per pipeline stage (one 128-byte swizzle row), taps x Cin/64 stages per tile.
// Warp roles (320 threads, persistent CTAs, static tile schedule): warp 0 TMA producer, warp 1 MMA issuer,
// warps 2-9 epilogue, two per TMEM lane quarter (TMEM -> registers -> scale/shift/residual/ReLU -> fp16 NDHWC stores).
// Two accumulator stages in TMEM (2 x BLOCK_N columns) overlap the epilogue of tile i with the MMAs of tile i+1.
// ConvTranspose(k=2,s=2) is the same GEMM with 4 (2D) / 8 (3D) weight groups stacked along N and a pixel-shuffle store.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "nrrt_common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // fp16 elements = 128 bytes = one SWIZZLE_128B row
constexpr int kEpiWarps = 8;                  // two warps per TMEM lane quarter
constexpr int kAuxWarp = 2 + kEpiWarps;        // warp 10: coefficient tiles of the polynomial-through-MMA kernels (else idle)
constexpr int kIssue2Warp = kAuxWarp + 1;      // warp 11: second MMA issuer of the pair slab kernel (else idle)
constexpr int kThreads = 64 + 32 * kEpiWarps + 64;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue, warps 10-11 as above
constexpr int kMaxTaps = 27;

struct ConvParams {
    int N, D, H, W;           // tile-space grid (conv: output grid; convT: input grid)
    int bw, bh, bd, rows;     // tile box, rows = bw*bh*bd <= 128
    int tiles_w, tiles_h, tiles_d, m_tiles;
    int stride;               // input coordinate = tile coordinate * stride + tap offset
    int stride_d;             // same for the depth axis (1 in 2D)
    int ntaps, cin_chunks, n_tiles;
    int split_chunk;          // K chunks [0, split) come from tensor map A, [split, cin_chunks) from tensor map A2
    const int32_t* idx1;      // optional sample indirection of source 1 / source 2 (encoder outputs shared per map)
    const int32_t* idx2;
    int8_t tap[kMaxTaps][4];  // (dx, dy, dz) including -pad
    // epilogue
    __half* out;
    long long out_cstride, out_coffset;
    int oD, oH, oW;           // output grid
    int up, up_d;             // 1 = same grid; 2 = ConvTranspose pixel shuffle (up_d for the depth axis)
    int cout;                 // channels per weight group (== total rows when up == 1)
    const float* scale;
    const float* shift;
    const float* scale2;  // optional second affine + ReLU
    const float* shift2;
    const __half* res;
    long long res_cstride, res_coffset;
    const int32_t* res_idx;   // optional sample indirection of the residual (per-map partial sums shared by the tasks of a map)
    int relu;
    const float* poly;   // optional [N][poly_cases][4][cout]: bilinear epilogue term E0 + ty*E1 + tx*E2 + ty*tx*E3
    int poly_cases;      // 9 = 3x3 border cases (row case * 3 + column case); else one case per weight group
    float inv_h, inv_w;  // 1/(H-1), 1/(W-1) of the tile-space grid
    int outer;           // slab kernels: slab loads per input-channel chunk: 3 (dx) in 2D, 9 (dz, dx) in 3D, 2 when folded
    int inner;           // slab kernels: vertical taps served by one slab: 3, or 2 when folded
    const float* poly3;  // pair slab kernel, 3D: [N][54][4][cout] coordinate field (see stage_poly3), applied before the ReLU
    float inv_d3, inv_h3;  // 1/(D-1), 1/(H-1)
    int dbg;             // experiments only (NRRT_CONV_DBG): 1 = no TMA stores, 2 = no epilogue arithmetic
    int res_k0;          // first K column of the identity block behind the weights (res_mma)
    int res_mma;         // pair slab kernel: the residual tile enters the accumulator through the tensor core (see below)
    int group;           // pair slab kernel: samples per scheduling group (tile-major inside a group), see halo_tile
    int fold;            // pair slab kernel: ConvTranspose(k=2,s=2) folded into the 3x3 conv that follows it (see conv_halo2_kernel)
    int ts;              // pair kernels: output tile staged in shared memory and written with TMA stores (see below)
    int bres;            // pair slab kernel: the layer's whole B (weights) fits the B ring exactly and stays RESIDENT: loaded
                         // during a CTA's first tile, never again (see conv_halo2_kernel)
    int pm;              // pair kernels: interior polynomial applied by one extra tcgen05.mma per tile (see pm_* below)
    const float* head_w; // optional fused 1x1 head (cout -> 1): logits[pixel] = head_b + sum_c relu_out[c] * head_w[c]
    float head_b;
    float* logits;
    uint32_t a_bytes, b_bytes;
};

// fp32 pair -> packed fp16x2 (lo in the low half) with the clamp to the finite fp16 range, and optionally the ReLU, inside
// the conversion instruction (F2FP.SATFINITE[.RELU]): relu + clamp + pack was 5 instructions per pair of outputs
__device__ __forceinline__ uint32_t pack_h2_sat(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t pack_h2_relu_sat(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}

// h-class of a row for the 3D coordinate field (csrc/conv_aux.cu, poly_relu3d_kernel): first row, inside piece 0, last row
// of piece 0, first row of piece 1, inside piece 1, last row.  Monotone in h.
__device__ __forceinline__ int hclass3(int h, int H) {
    const int hh = H >> 1;
    return h == 0 ? 0 : (h == H - 1 ? 5 : (h < hh - 1 ? 1 : (h == hh - 1 ? 2 : (h == hh ? 3 : 4))));
}

// Epilogue of 32 accumulator columns of one output pixel: folded-BN scale/shift, polynomial coordinate term, residual,
// ReLU, optional second affine + ReLU, then either the fp16 store into the output channel slice or (fused 1x1 head) a
// running dot product with the head weights.  col = first GEMM column, lc = its offset inside the staged scale/shift.
// RES_ST (pair kernels, compile time): res_row = this pixel's 128-byte row of the TMA-staged residual piece holding columns
// [lc & ~63, +64) (SWIZZLE_128B: 16-byte piece j of row r sits at j ^ (r & 7)); else the residual comes from global memory.
// OUT_ST: the fp16 results go to stg_row (same layout) for the TMA store; else to global memory.
// The epilogue warps share the SM's issue slots with nothing else that matters, and at full resolution a tile's epilogue
// (256 pixels x N columns) was ~430 instructions per 32 columns -- longer than the tile's MMAs for K <= 576 (ncu:
// profiles/conv_fold_r1.md).  Everything the staged paths do not need (pixel offsets, the weight-group division, global
// pointers) is compiled out of them, and ReLU / clamp ride on the conversion.
template <bool RES_ST, bool OUT_ST>
__device__ __forceinline__ void epilogue_chunk(const ConvParams& p, const uint32_t (&v)[32], int col, int lc, int n, int d,
                                               int h, int w, const float* sc, const float* sh, const float* sc2,
                                               const float* sh2, const float* pe, int pe_stride, float& head_acc,
                                               const uint8_t* res_row = nullptr, int res_r7 = 0, bool pm = false,
                                               int frame = 0, uint8_t* stg_row = nullptr, int slot3 = 0) {
    int g = 0, co = col;
    if (p.up == 2) { g = col / p.cout; co = col - g * p.cout; }
    auto out_pixel = [&](int ns) -> long long {  // pixel index in the output grid (ConvTranspose: pixel shuffle by group)
        int od = d, oh = h, ow = w;
        if (p.up == 2) {
            if (p.up_d == 2) { od = 2 * d + (g >> 2); oh = 2 * h + ((g >> 1) & 1); ow = 2 * w + (g & 1); }
            else { oh = 2 * h + (g >> 1); ow = 2 * w + (g & 1); }
        }
        return (((long long)ns * p.oD + od) * p.oH + oh) * p.oW + ow;
    };
    float f[32];
    if (p.scale) {
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = fmaf(__uint_as_float(v[i]), sc[lc + i], sh[lc + i]);
    } else {  // identity affine (scale == NULL): the layer's bias rides in the polynomial's constant term
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
    }
    if (p.poly) {
        int kcase = g;
        if (p.poly_cases == 9)
            kcase = (h == 0 ? 0 : (h == p.H - 1 ? 2 : 1)) * 3 + (w == 0 ? 0 : (w == p.W - 1 ? 2 : 1));
        if (!pm) {
            // interior pixels (and ConvTranspose groups) read the tile's coefficients staged in shared memory; only the
            // one-pixel image frame, whose tap set differs, goes to global memory
            const float ty = (float)h * p.inv_h, tx = (float)w * p.inv_w, tyx = ty * tx;
            const bool staged = p.poly_cases != 9 || kcase == 4;
            const float4* e0 = staged ? reinterpret_cast<const float4*>(pe + lc)
                                      : reinterpret_cast<const float4*>(p.poly + (((size_t)n * p.poly_cases + kcase) * 4) * p.cout + co);
            const int es = staged ? pe_stride / 4 : p.cout / 4;
            const float4* e1 = e0 + es;
            const float4* e2 = e1 + es;
            const float4* e3 = e2 + es;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 a0 = e0[q], a1 = e1[q], a2 = e2[q], a3 = e3[q];
                f[q * 4 + 0] += a0.x + ty * a1.x + tx * a2.x + tyx * a3.x;
                f[q * 4 + 1] += a0.y + ty * a1.y + tx * a2.y + tyx * a3.y;
                f[q * 4 + 2] += a0.z + ty * a1.z + tx * a2.z + tyx * a3.z;
                f[q * 4 + 3] += a0.w + ty * a1.w + tx * a2.w + tyx * a3.w;
            }
        } else if (frame) {
            // the extra MMA added the interior polynomial to every pixel: frame pixels get (their case - interior), staged per
            // tile by pm_stage_frame in rows [A_r, B_r, A_c, B_c, C] behind scale/shift: row frame A_r + tx*B_r, column
            // frame A_c + ty*B_c, corner C
            const float* fr = sc2 + lc;  // == tile buffer + 2 * BLOCK_N + lc
            const float* a = frame == 1 ? fr : (frame == 2 ? fr + 2 * pe_stride : fr + 4 * pe_stride);
            const float t = frame == 1 ? (float)w * p.inv_w : (frame == 2 ? (float)h * p.inv_h : 0.f);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 a0 = *reinterpret_cast<const float4*>(a + q * 4);
                const float4 a1 = frame == 3 ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(a + pe_stride + q * 4);
                f[q * 4 + 0] += a0.x + t * a1.x;
                f[q * 4 + 1] += a0.y + t * a1.y;
                f[q * 4 + 2] += a0.z + t * a1.z;
                f[q * 4 + 3] += a0.w + t * a1.w;
            }
        }
    }
    if (p.poly3) {
        // 3D coordinate field of the tile's depth slice: A + th * B per (h-class, channel), staged by stage_poly3 for the
        // interior w case; the two frame columns (w = 0, W - 1) have their own tap set and read their case from global memory
        const float th = (float)h * p.inv_h3;
        if (w != 0 && w != p.W - 1) {
            const float4* A = reinterpret_cast<const float4*>(sc + (2 + 2 * slot3) * pe_stride + lc);
            const float4* Bq = reinterpret_cast<const float4*>(sc + (3 + 2 * slot3) * pe_stride + lc);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 a = A[q], b = Bq[q];
                f[q * 4 + 0] += fmaf(th, b.x, a.x); f[q * 4 + 1] += fmaf(th, b.y, a.y);
                f[q * 4 + 2] += fmaf(th, b.z, a.z); f[q * 4 + 3] += fmaf(th, b.w, a.w);
            }
        } else {
            const int dc = d == 0 ? 0 : (d == p.D - 1 ? 2 : 1);
            const float td = (float)d * p.inv_d3;
            const float* e = p.poly3 + (((size_t)n * 54 + (size_t)(dc * 6 + hclass3(h, p.H)) * 3 + (w == 0 ? 0 : 2)) * 4) * p.cout + co;
#pragma unroll
            for (int i = 0; i < 32; ++i)
                f[i] += __ldg(e + i) + td * __ldg(e + p.cout + i) + th * (__ldg(e + 2 * (size_t)p.cout + i) + td * __ldg(e + 3 * (size_t)p.cout + i));
        }
    }
    const int j0 = (lc & 63) >> 3;  // first 16-byte piece of these 32 columns inside a staged 128-byte row
    if (RES_ST || p.res) {
        const uint4* rp = nullptr;
        if (!RES_ST) rp = reinterpret_cast<const uint4*>(p.res + out_pixel(p.res_idx ? __ldg(p.res_idx + n) : n) * p.res_cstride +
                                                         p.res_coffset + co);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 rv = RES_ST ? *reinterpret_cast<const uint4*>(res_row + (((j0 + q) ^ res_r7) << 4)) : __ldg(rp + q);
            const __half2* hv = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 t2 = __half22float2(hv[e]);
                f[q * 8 + e * 2] += t2.x;
                f[q * 8 + e * 2 + 1] += t2.y;
            }
        }
    }
    if (p.logits || p.scale2) {  // the ReLU as arithmetic only where a float consumer follows; else it rides on the conversion
        if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.0f);
        }
        if (p.scale2) {
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = fmaxf(fmaf(f[i], sc2[lc + i], sh2[lc + i]), 0.0f);
        }
        if (p.logits) {  // fused final_conv: the 64-channel activation never goes to memory
#pragma unroll
            for (int i = 0; i < 32; ++i) head_acc = fmaf(f[i], __ldg(p.head_w + co + i), head_acc);
            return;
        }
    }
    uint4 ov[4];
    if (p.relu) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            ov[q].x = pack_h2_relu_sat(f[q * 8 + 0], f[q * 8 + 1]); ov[q].y = pack_h2_relu_sat(f[q * 8 + 2], f[q * 8 + 3]);
            ov[q].z = pack_h2_relu_sat(f[q * 8 + 4], f[q * 8 + 5]); ov[q].w = pack_h2_relu_sat(f[q * 8 + 6], f[q * 8 + 7]);
        }
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            ov[q].x = pack_h2_sat(f[q * 8 + 0], f[q * 8 + 1]); ov[q].y = pack_h2_sat(f[q * 8 + 2], f[q * 8 + 3]);
            ov[q].z = pack_h2_sat(f[q * 8 + 4], f[q * 8 + 5]); ov[q].w = pack_h2_sat(f[q * 8 + 6], f[q * 8 + 7]);
        }
    }
    if (OUT_ST) {  // staged for the TMA store
#pragma unroll
        for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(stg_row + (((j0 + q) ^ res_r7) << 4)) = ov[q];
    } else {
        uint4* op = reinterpret_cast<uint4*>(p.out + out_pixel(n) * p.out_cstride + p.out_coffset + co);
#pragma unroll
        for (int q = 0; q < 4; ++q) op[q] = ov[q];
    }
}

// Per-tile staging by the 256 epilogue threads: scale/shift (+ second affine) and the tile's polynomial coefficients.
template <int BLOCK_N>
__device__ __forceinline__ void stage_tile_params(const ConvParams& p, float* dst, int col0, int n, int et, bool affine,
                                                  bool poly) {
    affine = affine && p.scale;
    if (!affine && !(poly && p.poly)) return;  // nothing changed since this buffer was last filled (uniform per CTA)
    // every epilogue warp must be done reading the old contents (barriers are skipped on tiles that restage nothing)
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
    if (affine)
    for (int i = et; i < BLOCK_N; i += 32 * kEpiWarps) {
        dst[i] = __ldg(p.scale + col0 + i);
        dst[BLOCK_N + i] = __ldg(p.shift + col0 + i);
        if (p.scale2) {
            dst[2 * BLOCK_N + i] = __ldg(p.scale2 + col0 + i);
            dst[3 * BLOCK_N + i] = __ldg(p.shift2 + col0 + i);
        }
    }
    if (p.poly && poly) {
        const int g = col0 / p.cout, co0 = col0 - g * p.cout;
        const int kcase = p.poly_cases == 9 ? 4 : g;
        const float* src = p.poly + (((size_t)n * p.poly_cases + kcase) * 4) * p.cout + co0;
        for (int i = et; i < 4 * BLOCK_N; i += 32 * kEpiWarps) {
            const int e = i / BLOCK_N, c = i - e * BLOCK_N;
            if (co0 + c < p.cout) dst[(4 + e) * BLOCK_N + c] = __ldg(src + (size_t)e * p.cout + c);
        }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
}

// PM kernels: per-tile frame corrections (see epilogue_chunk) for a tile that touches the one-pixel image frame, written by
// the 256 epilogue threads into rows 2..6 of the tile's parameter buffer.  hr / wc = tile-local index of the frame row /
// column (-1: none), rr / cc = its border case (0 = first, 2 = last).  Tile-uniform: every epilogue thread calls it.
template <int BLOCK_N>
__device__ __forceinline__ void pm_stage_frame(const ConvParams& p, float* dst, int n, int et, int h0, int w0, int hr, int wc,
                                               int col0 = 0) {
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
    const int rr = hr < 0 ? 1 : (h0 + hr == 0 ? 0 : 2), cc = wc < 0 ? 1 : (w0 + wc == 0 ? 0 : 2);
    const float ty_r = (float)(h0 + hr) * p.inv_h, tx_c = (float)(w0 + wc) * p.inv_w;
    const float* base = p.poly + ((size_t)n * 9 * 4) * p.cout + col0;
    for (int i = et; i < 3 * BLOCK_N; i += 32 * kEpiWarps) {
        const int which = i / BLOCK_N, c = i - which * BLOCK_N;  // 0: row frame, 1: column frame, 2: corner
        if ((which == 0 && hr < 0) || (which == 1 && wc < 0) || (which == 2 && (hr < 0 || wc < 0))) continue;
        const int kcase = which == 0 ? rr * 3 + 1 : (which == 1 ? 3 + cc : rr * 3 + cc);
        const float* k = base + (size_t)kcase * 4 * p.cout + c;
        const float* q = base + (size_t)4 * 4 * p.cout + c;
        const float d0 = __ldg(k) - __ldg(q), d1 = __ldg(k + p.cout) - __ldg(q + p.cout);
        const float d2 = __ldg(k + 2 * (size_t)p.cout) - __ldg(q + 2 * (size_t)p.cout);
        const float d3 = __ldg(k + 3 * (size_t)p.cout) - __ldg(q + 3 * (size_t)p.cout);
        if (which == 0) {
            dst[2 * BLOCK_N + c] = d0 + ty_r * d1;
            dst[3 * BLOCK_N + c] = d2 + ty_r * d3;
        } else if (which == 1) {
            dst[4 * BLOCK_N + c] = d0 + tx_c * d2;
            dst[5 * BLOCK_N + c] = d1 + tx_c * d3;
        } else {
            dst[6 * BLOCK_N + c] = d0 + ty_r * d1 + tx_c * d2 + ty_r * tx_c * d3;
        }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
}

// 3D coordinate field in the epilogue (conv8.0 of ThreeD_UNet_cooler; ThreeD_train.py:180-191, derivation in conv_aux.cu).
// A tile is one depth slice, so td is tile-uniform and the field E0 + td E1 + th (E2 + td E3) is A + th * B per (h-class,
// channel).  The 256 epilogue threads stage A, B of the (up to 6) h-classes of the tile's rows for the interior w case into
// rows 2 + 2 q, 3 + 2 q of the tile's parameter buffer (16 rows per accumulator stage in these kernels), q = class - class
// of the tile's first row.  Replaces the read-modify-write post-pass over the widest activation of the 3D decoder
// (nrrt_poly_relu_3d: 655 MB in + 655 MB out per 5 tasks).
template <int BLOCK_N>
__device__ __forceinline__ void stage_poly3(const ConvParams& p, float* dst, int n, int td, int h0, int et, int col0) {
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
    const int dc = td == 0 ? 0 : (td == p.D - 1 ? 2 : 1);
    const float tdf = (float)td * p.inv_d3;
    const int hc0 = hclass3(h0, p.H), hc1 = hclass3(min(h0 + 15, p.H - 1), p.H);
    for (int i = et; i < (hc1 - hc0 + 1) * BLOCK_N; i += 32 * kEpiWarps) {
        const int q = i / BLOCK_N, c = i - q * BLOCK_N;
        const float* e = p.poly3 + (((size_t)n * 54 + (size_t)(dc * 6 + hc0 + q) * 3 + 1) * 4) * p.cout + col0 + c;
        const float e0 = __ldg(e), e1 = __ldg(e + p.cout), e2 = __ldg(e + 2 * (size_t)p.cout), e3 = __ldg(e + 3 * (size_t)p.cout);
        dst[(2 + 2 * q) * BLOCK_N + c] = e0 + tdf * e1;
        dst[(3 + 2 * q) * BLOCK_N + c] = e2 + tdf * e3;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
}

template <int BLOCK_N, int STAGES>
struct SmemLayout {
    static constexpr uint32_t a_stage = kBlockM * 128;
    static constexpr uint32_t b_stage = BLOCK_N * 128;
    static constexpr uint32_t off_b = STAGES * a_stage;
    static constexpr uint32_t off_bar = off_b + STAGES * b_stage;
    static constexpr uint32_t off_tmem = off_bar + (2 * STAGES + 4) * 8;
    static constexpr uint32_t off_ss = (off_tmem + 16 + 15) & ~15u;         // [2 acc stages][scale, shift, scale2, shift2, E0..E3][BLOCK_N] floats
    static constexpr uint32_t total = off_ss + 2 * 8 * BLOCK_N * 4 + 1024;  // + slack for manual 1024-B alignment
};

template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2,
                 const __grid_constant__ CUtensorMap map_b, const __grid_constant__ ConvParams p) {
    using L = SmemLayout<BLOCK_N, STAGES>;
    constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar0 = base + L::off_bar;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 2 + s); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + L::off_tmem);
    float* ss = reinterpret_cast<float*>(base_ptr + L::off_ss);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.m_tiles * p.n_tiles;
    const int k_iters = p.ntaps * p.cin_chunks;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + L::off_tmem),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int n_tile = tile % p.n_tiles;
                int m = tile / p.n_tiles;
                const int tw = m % p.tiles_w; m /= p.tiles_w;
                const int th = m % p.tiles_h; m /= p.tiles_h;
                const int td = m % p.tiles_d;
                const int n = m / p.tiles_d;
                const int w0 = tw * p.bw * p.stride, h0 = th * p.bh * p.stride, d0 = td * p.bd * p.stride_d;
                const int n1 = p.idx1 ? __ldg(p.idx1 + n) : n, n2 = p.idx2 ? __ldg(p.idx2 + n) : n;
                for (int t = 0; t < p.ntaps; ++t) {
                    const int cw = w0 + p.tap[t][0], ch = h0 + p.tap[t][1], cd = d0 + p.tap[t][2];
                    for (int c = 0; c < p.cin_chunks; ++c) {
                        mbar_wait(empty_bar(stage), phase ^ 1u);
                        mbar_expect_tx(full_bar(stage), p.a_bytes + p.b_bytes);
                        if (c < p.split_chunk)
                            tma_load_5d(base + stage * L::a_stage, &map_a, full_bar(stage), c * kBlockK, cw, ch, cd, n1);
                        else
                            tma_load_5d(base + stage * L::a_stage, &map_a2, full_bar(stage), (c - p.split_chunk) * kBlockK, cw, ch, cd, n2);
                        tma_load_2d(base + L::off_b + stage * L::b_stage, &map_b, full_bar(stage),
                                    (t * p.cin_chunks + c) * kBlockK, n_tile * BLOCK_N);
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected thread) =====
        {  // the whole warp runs the loop (uniform control flow: descriptors stay in uniform registers); one lane issues
            constexpr uint32_t idesc = make_idesc(BLOCK_N);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int k = 0; k < k_iters; ++k) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc(base + stage * L::a_stage);
                    const uint64_t bdesc = make_smem_desc(base + L::off_b + stage * L::b_stage);
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < kBlockK / 16; ++kk)  // UMMA_K = 16: +32 bytes inside the swizzle row
                            tc_mma_f16(d_tmem, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (k | kk) ? 1u : 0u);
                        tc_commit(empty_bar(stage));  // frees the smem slot once these MMAs have read it
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                if (elect_one()) tc_commit(tfull_bar(acc));  // accumulator complete
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else if (warp < kAuxWarp) {
        // ===== epilogue warps 2..9: TMEM lanes 32*(warp%4) .. +31; the two warps of a quarter split the column chunks =====
        const int sub = warp & 3;
        const int part = (warp - 2) >> 2;
        const int row = sub * 32 + lane;
        const int et = threadIdx.x - 64;  // 0..255
        int acc = 0;
        uint32_t acc_phase = 0;
        int last_ntile[2] = {-1, -1}, last_n[2] = {-1, -1};
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int n_tile = tile % p.n_tiles;
            int m = tile / p.n_tiles;
            const int tw = m % p.tiles_w; m /= p.tiles_w;
            const int th = m % p.tiles_h; m /= p.tiles_h;
            const int td = m % p.tiles_d;
            const int n = m / p.tiles_d;
            float* sc = ss + acc * 8 * BLOCK_N;  // double-buffered by accumulator stage
            stage_tile_params<BLOCK_N>(p, sc, n_tile * BLOCK_N, n, et, last_ntile[acc] != n_tile,
                                       last_ntile[acc] != n_tile || last_n[acc] != n);
            last_ntile[acc] = n_tile; last_n[acc] = n;

            const int wl = row % p.bw, hl = (row / p.bw) % p.bh, dl = row / (p.bw * p.bh);
            const int w = tw * p.bw + wl, h = th * p.bh + hl, d = td * p.bd + dl;
            const bool valid = row < p.rows && w < p.W && h < p.H && d < p.D;

            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            float head_acc = 0.f;
            // fused head: one warp per quarter walks all chunks (the dot product runs over all channels of a pixel)
            const int c_begin = p.logits ? 0 : part, c_step = p.logits ? 1 : 2;
            if (!(p.logits && part == 1)) {
#pragma unroll 1
                for (int c = c_begin; c < BLOCK_N / 32; c += c_step) {
                    uint32_t v[32];
                    tc_ld32(taddr + (uint32_t)(c * 32), v);
                    if (valid)
                        epilogue_chunk<false, false>(p, v, n_tile * BLOCK_N + c * 32, c * 32, n, d, h, w, sc, sc + BLOCK_N, sc + 2 * BLOCK_N,
                                       sc + 3 * BLOCK_N, sc + 4 * BLOCK_N, BLOCK_N, head_acc);
                }
                if (valid && p.logits) p.logits[(((long long)n * p.oD + d) * p.oH + h) * p.oW + w] = head_acc + p.head_b;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}


// ---- cta_group::2 (CTA pair) primitives ----------------------------------------------------------------------------
// In a 2-CTA cluster the shared::cluster addresses of the two CTAs differ in bit 24; clearing it addresses the leader.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma2_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2,
                                             int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::
            "r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_mma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
// commit of the pair's MMAs, arriving on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void tc_commit2(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
// arrive on the leader CTA's copy of a barrier (local for the leader, remote for the peer)
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar)
        : "memory");
}

constexpr int kMaxResPieces = 4;       // 64-channel pieces of a residual tile (BLOCK_N <= 256)
constexpr uint32_t kResPiece = 16384;  // 128 pixel rows x 128 B (SWIZZLE_128B), 1024-byte aligned


// ---- polynomial coordinate term through the tensor core (pair kernels, 3x3 convs) ---------------------------------
// The epilogue polynomial E0 + ty*E1 + tx*E2 + ty*tx*E3 (4 coefficient loads + 4 FMAs per output value) is the largest
// single cost of the conv7.0 / conv8.0 epilogues.  Inside one tile it is a rank-4 product: with tile-local pixel
// coordinates (hl, wl) in [0,16) it equals [1, hl, wl, hl*wl] . [E0', E1', E2', E3'] where the primed coefficients
// absorb the tile origin (pm_produce_b).  The basis is exact in fp16 and the SAME for every tile, so it is a constant
// shared-memory A tile written once per CTA; the coefficients are split hi + lo into two fp16 values (22 significant
// bits) and written per tile as a K = 16 B tile: one extra K=16 MMA per accumulator replaces the whole interior
// polynomial.  A CTA pair shares B, so the 16 K slots are [tile of CTA 0: 4 hi, 4 lo | tile of CTA 1: 4 hi, 4 lo] and
// each CTA's basis rows are zero in the other CTA's slots.  Pixels on the image frame (different tap set) get the
// difference to the interior polynomial in the epilogue.  Both tiles use the no-swizzle K-major canonical layout
// ((8,n),2):((16 B, SBO), LBO): [k half][row][8 halfs], SBO = 128 B, LBO = rows * 16 B.
__device__ __forceinline__ uint64_t make_smem_desc_nosw(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(lbo >> 4) << 16;
    d |= (uint64_t)(sbo >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;  // layout type 0: no swizzle
}
#define mbar_wait_cluster(bar, parity)                                                    \
    asm volatile(                                                                         \
        "{\n\t.reg .pred p;\n\t"                                                         \
        "WAITC_LOOP:\n\t"                                                                 \
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"       \
        "@p bra WAITC_DONE;\n\t"                                                          \
        "bra WAITC_LOOP;\n\t"                                                             \
        "WAITC_DONE:\n\t}" ::"r"((uint32_t)(bar)), "r"((uint32_t)(parity))                \
        : "memory")
__device__ __forceinline__ void mbar_arrive_leader_release(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar)
        : "memory");
}
// basis rows of one 128-pixel block: row r holds [1, hl, wl, hl*wl, 1, hl, wl, hl*wl] in k half `rank`, zeros in the other
__device__ __forceinline__ void pm_write_basis(uint8_t* tile, int row, int hl, int wl, uint32_t rank) {
    const __half2 a = __floats2half2_rn(1.0f, (float)hl), b = __floats2half2_rn((float)wl, (float)(hl * wl));
    uint4 v;
    v.x = *reinterpret_cast<const uint32_t*>(&a); v.y = *reinterpret_cast<const uint32_t*>(&b); v.z = v.x; v.w = v.y;
    *reinterpret_cast<uint4*>(tile + rank * 2048u + row * 16) = v;
    *reinterpret_cast<uint4*>(tile + (1u - rank) * 2048u + row * 16) = make_uint4(0u, 0u, 0u, 0u);
}
// B tile of one pair iteration: rows = this CTA's ROWS_B output channels, k half j = tile of CTA j (sample tn[j], origin
// (th0[j], tw0[j]) in pixels, dead tiles contribute zero).  Run by the auxiliary warp (thread t0 of nt), off the critical
// path of the epilogue, the TMA producer and the MMA issuer; ends with the writes visible to the async proxy.
template <int ROWS_B>
__device__ __forceinline__ void pm_produce_b(const ConvParams& p, uint8_t* bx, int t0, int nt, uint32_t rank, const int (&tn)[2],
                                             const int (&th0)[2], const int (&tw0)[2], const bool (&tlive)[2], int col0 = 0) {
    constexpr int kMaxItems = (2 * ROWS_B + 31) / 32;  // items per thread of the auxiliary warp
    float e[kMaxItems][4];
    // all coefficient loads first (independent, one latency), then the arithmetic
#pragma unroll
    for (int i = 0; i < kMaxItems; ++i) {
        const int it = t0 + i * nt;
        const int j = it / ROWS_B, r = it - j * ROWS_B;
        const int co = col0 + (int)rank * ROWS_B + r;
        e[i][0] = e[i][1] = e[i][2] = e[i][3] = 0.f;
        if (it < 2 * ROWS_B && tlive[j] && co < p.cout) {
            const float* src = p.poly + (((size_t)tn[j] * 9 + 4) * 4) * p.cout + co;  // interior case
            e[i][0] = __ldg(src); e[i][1] = __ldg(src + p.cout);
            e[i][2] = __ldg(src + 2 * (size_t)p.cout); e[i][3] = __ldg(src + 3 * (size_t)p.cout);
        }
    }
#pragma unroll
    for (int i = 0; i < kMaxItems; ++i) {
        const int it = t0 + i * nt;
        if (it >= 2 * ROWS_B) break;
        const int j = it / ROWS_B, r = it - j * ROWS_B;
        const float e0 = e[i][0], e1 = e[i][1], e2 = e[i][2], e3 = e[i][3];
        const float hy = (float)th0[j] * p.inv_h, wx = (float)tw0[j] * p.inv_w;  // ty, tx of the tile origin
        const float c0 = e0 + hy * e1 + wx * e2 + hy * wx * e3;
        const float c1 = p.inv_h * (e1 + wx * e3);
        const float c2 = p.inv_w * (e2 + hy * e3);
        const float c3 = p.inv_h * p.inv_w * e3;
        const float c[4] = {c0, c1, c2, c3};
        __half hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float cc = fminf(fmaxf(c[q], -65504.f), 65504.f);
            hi[q] = __float2half_rn(cc);
            lo[q] = __float2half_rn(cc - __half2float(hi[q]));
        }
        uint4 v;
        __half2 t;
        t = __halves2half2(hi[0], hi[1]); v.x = *reinterpret_cast<const uint32_t*>(&t);
        t = __halves2half2(hi[2], hi[3]); v.y = *reinterpret_cast<const uint32_t*>(&t);
        t = __halves2half2(lo[0], lo[1]); v.z = *reinterpret_cast<const uint32_t*>(&t);
        t = __halves2half2(lo[2], lo[3]); v.w = *reinterpret_cast<const uint32_t*>(&t);
        *reinterpret_cast<uint4*>(bx + (uint32_t)j * (ROWS_B * 16u) + r * 16) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

template <int BLOCK_N, int STAGES, bool RES, bool PM, bool TS>
struct SmemLayout2 {  // per CTA of a pair: its own A tile and HALF of the B tile per stage
    static constexpr uint32_t a_stage = kBlockM * 128;
    static constexpr uint32_t b_stage = (BLOCK_N / 2) * 128;
    static constexpr uint32_t off_b = STAGES * a_stage;
    static constexpr uint32_t off_res = off_b + STAGES * b_stage;  // TMA-staged residual tile: BLOCK_N / 64 pieces
    static constexpr uint32_t off_pm = off_res + ((RES || TS) ? (BLOCK_N / 64) * kResPiece : 0);  // basis tile (4 KB), then 2 coefficient tiles
    static constexpr uint32_t pm_b = (BLOCK_N / 2) * 32;
    static constexpr uint32_t off_bar = off_pm + (PM ? 4096 + 2 * pm_b : 0);
    static constexpr uint32_t off_tmem = off_bar + (2 * STAGES + 4 + 3 * kMaxResPieces + 2) * 8;
    static constexpr uint32_t off_ss = (off_tmem + 16 + 15) & ~15u;
    static constexpr uint32_t total = off_ss + 2 * 8 * BLOCK_N * 4 + 1024;
};

// conv_gemm2_kernel — conv_gemm_kernel on CTA pairs (tcgen05 cta_group::2).  tcgen05.mma in SS mode is fed from shared
// memory at ~85 B/clk/SM (profiles/README.md), which caps a 1-CTA M128xN256 MMA at ~86 % and narrower N lower.  A CTA pair
// issues M256 MMAs: each SM supplies its own 128 A rows and only HALF of B (the other half comes from the peer's shared
// memory), cutting operand bytes per SM from 4 KB + 32N to 4 KB + 16N per K=16 step.  Both CTAs run a TMA producer (own A
// tile + own B half, all signalling the leader's barrier) and the epilogue; only the leader issues MMAs and its commits
// are multicast to the barriers of both CTAs.
// RES: the residual tile (this CTA's 128 pixels x BLOCK_N channels, fp16) is staged in shared memory by the producer's
// TMA in 64-channel pieces, each with a full/empty barrier pair, so the epilogue never waits on a global load: the piece
// of tile i+1 is requested as soon as every epilogue warp has consumed that piece of tile i.
// PM: the interior polynomial coordinate term is one extra MMA per tile (pm_* above; 2D 3x3 layers, n_tiles == 1).
// TS: TMA-store epilogue (64-channel pieces shared with the residual staging; the auxiliary warp issues the stores).
template <int BLOCK_N, int STAGES, bool RES, bool PM, bool TS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_gemm2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2,
                 const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_r,
                 const __grid_constant__ CUtensorMap map_o, const __grid_constant__ ConvParams p) {
    using L = SmemLayout2<BLOCK_N, STAGES, RES, PM, TS>;
    constexpr int RP = BLOCK_N / 64;  // residual pieces
    constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar0 = base + L::off_bar;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 2 + s); };
    auto rfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 4 + s); };
    auto rempty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 4 + kMaxResPieces + s); };
    auto sdone_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 4 + 2 * kMaxResPieces + s); };
    auto bxfull_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + 4 + 3 * kMaxResPieces + s); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + L::off_tmem);
    float* ss = reinterpret_cast<float*>(base_ptr + L::off_ss);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();       // 0 = leader (issues the MMAs for the pair), 1 = peer
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int total_pairs = ((p.m_tiles + 1) >> 1) * p.n_tiles;  // a pair of CTAs owns two consecutive M tiles
    const int k_iters = p.ntaps * p.cin_chunks;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        // the leader's tempty barrier collects the epilogue warps of BOTH CTAs
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 2 * kEpiWarps); }
        if (RES || TS)  // piece free: all epilogue warps, or (TS) the store warp once the piece has been read out
            for (int s = 0; s < RP; ++s) {
                mbar_init(rfull_bar(s), 1); mbar_init(rempty_bar(s), TS ? 1 : kEpiWarps); mbar_init(sdone_bar(s), kEpiWarps);
            }
        if (PM) for (int s = 0; s < 2; ++s) mbar_init(bxfull_bar(s), 2);  // one arrival per CTA of the pair
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + L::off_tmem),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // barriers of both CTAs are initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0, rphase = 0;
            for (int pt = cluster_id; pt < total_pairs; pt += n_clusters) {
                const int n_tile = pt % p.n_tiles;
                int m = (pt / p.n_tiles) * 2 + (int)rank;
                const bool live = m < p.m_tiles;  // an odd tile count leaves the peer of the last pair without work
                const int tw = m % p.tiles_w; m /= p.tiles_w;
                const int th = m % p.tiles_h; m /= p.tiles_h;
                const int td = m % p.tiles_d;
                const int n = m / p.tiles_d;
                const int w0 = tw * p.bw * p.stride, h0 = th * p.bh * p.stride, d0 = td * p.bd * p.stride_d;
                // a dead tile still loads (zero-filled: sample index past the end) so the pair's byte count is constant
                const int n1 = !live ? 0x3fffffff : (p.idx1 ? __ldg(p.idx1 + n) : n);
                const int n2 = !live ? 0x3fffffff : (p.idx2 ? __ldg(p.idx2 + n) : n);
                for (int t = 0; t < p.ntaps; ++t) {
                    const int cw = w0 + p.tap[t][0], ch = h0 + p.tap[t][1], cd = d0 + p.tap[t][2];
                    for (int c = 0; c < p.cin_chunks; ++c) {
                        mbar_wait(empty_bar(stage), phase ^ 1u);
                        // one arrival (the leader's) expects the bytes of both CTAs; every load signals the leader's barrier
                        if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * p.a_bytes + p.b_bytes);
                        const uint32_t lbar = full_bar(stage) & kPeerBitMask;
                        if (c < p.split_chunk)
                            tma2_load_5d(base + stage * L::a_stage, &map_a, lbar, c * kBlockK, cw, ch, cd, n1);
                        else
                            tma2_load_5d(base + stage * L::a_stage, &map_a2, lbar, (c - p.split_chunk) * kBlockK, cw, ch, cd, n2);
                        tma2_load_2d(base + L::off_b + stage * L::b_stage, &map_b, lbar, (t * p.cin_chunks + c) * kBlockK,
                                     n_tile * BLOCK_N + (int)rank * (BLOCK_N / 2));
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
                if (RES) {
                    // residual of this tile (own CTA, own barriers), issued behind the tile's operand loads
                    const int nr = !live ? 0x3fffffff : (p.res_idx ? __ldg(p.res_idx + n) : n);
                    for (int s = 0; s < RP; ++s) {
                        mbar_wait(rempty_bar(s), rphase ^ 1u);
                        mbar_expect_tx(rfull_bar(s), p.a_bytes);
                        tma_load_5d(base + L::off_res + s * kResPiece, &map_r, rfull_bar(s),
                                    (int)p.res_coffset + n_tile * BLOCK_N + s * 64, tw * p.bw, th * p.bh, td * p.bd, nr);
                    }
                    rphase ^= 1u;
                }
            }
            // drain: the leader's multicast commits must not land in a CTA that already exited
            for (int i = 0; i < STAGES; ++i) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one thread of the leader CTA drives both SMs' tensor cores (M = 256 per instruction) =====
        if (rank == 0) {  // the whole warp runs the loop (uniform control flow: descriptors stay in uniform registers)
            constexpr uint32_t idesc = make_idesc(BLOCK_N, 256);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int pt = cluster_id; pt < total_pairs; pt += n_clusters) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int k = 0; k < k_iters; ++k) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc(base + stage * L::a_stage);
                    const uint64_t bdesc = make_smem_desc(base + L::off_b + stage * L::b_stage);
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < kBlockK / 16; ++kk)  // UMMA_K = 16: +32 bytes inside the swizzle row
                            tc_mma2_f16(d_tmem, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc, (k | kk) ? 1u : 0u);
                        tc_commit2(empty_bar(stage));  // frees the smem slot in both CTAs once these MMAs have read it
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                if (PM) {
                    mbar_wait_cluster(bxfull_bar(acc), acc_phase);  // both CTAs' coefficient tiles of this pair iteration
                    tc_fence_after();
                    if (elect_one())
                        tc_mma2_f16(d_tmem, make_smem_desc_nosw(base + L::off_pm, 2048, 128),
                                   make_smem_desc_nosw(base + L::off_pm + 4096 + acc * L::pm_b, (BLOCK_N / 2) * 16, 128), idesc, 1u);
                    __syncwarp();
                }
                if (elect_one()) tc_commit2(tfull_bar(acc));  // accumulator complete in both CTAs' tensor memory
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else if (warp < kAuxWarp) {
        // ===== epilogue warps 2..9: TMEM lanes 32*(warp%4) .. +31; the two warps of a quarter split the column chunks =====
        const int sub = warp & 3;
        const int part = (warp - 2) >> 2;
        const int row = sub * 32 + lane;
        const int et = threadIdx.x - 64;  // 0..255
        int acc = 0;
        uint32_t acc_phase = 0, rphase = 0;
        int last_ntile[2] = {-1, -1}, last_n[2] = {-1, -1};
        for (int pt = cluster_id; pt < total_pairs; pt += n_clusters) {
            const int n_tile = pt % p.n_tiles;
            int m = (pt / p.n_tiles) * 2 + (int)rank;
            const bool live = m < p.m_tiles;
            const int tw = m % p.tiles_w; m /= p.tiles_w;
            const int th = m % p.tiles_h; m /= p.tiles_h;
            const int td = m % p.tiles_d;
            const int n = live ? m / p.tiles_d : 0;
            float* sc = ss + acc * 8 * BLOCK_N;  // double-buffered by accumulator stage
            stage_tile_params<BLOCK_N>(p, sc, n_tile * BLOCK_N, n, et, last_ntile[acc] != n_tile,
                                       !PM && (last_ntile[acc] != n_tile || last_n[acc] != n));
            last_ntile[acc] = n_tile; last_n[acc] = n;

            const int wl = row % p.bw, hl = (row / p.bw) % p.bh, dl = row / (p.bw * p.bh);
            const int w = tw * p.bw + wl, h = th * p.bh + hl, d = td * p.bd + dl;
            const bool valid = live && row < p.rows && w < p.W && h < p.H && d < p.D;
            int frame = 0;
            if (PM) {
                const int h0 = th * p.bh, w0 = tw * p.bw;
                const int hr = h0 == 0 ? 0 : (p.H - 1 - h0 < p.bh ? p.H - 1 - h0 : -1);
                const int wc = w0 == 0 ? 0 : (p.W - 1 - w0 < p.bw ? p.W - 1 - w0 : -1);
                if (live && (hr >= 0 || wc >= 0)) {
                    pm_stage_frame<BLOCK_N>(p, sc, n, et, h0, w0, hr, wc);
                    const bool on_r = hl == hr, on_c = wl == wc;
                    frame = on_r ? (on_c ? 3 : 1) : (on_c ? 2 : 0);
                }
            }

            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * BLOCK_N);
            float head_acc = 0.f;
            // fused head: one warp per quarter walks all chunks (the dot product runs over all channels of a pixel)
            const int c_begin = p.logits ? 0 : part, c_step = p.logits ? 1 : 2;
            if (!(p.logits && part == 1)) {
#pragma unroll 1
                for (int c = c_begin; c < BLOCK_N / 32; c += c_step) {
                    uint32_t v[32];
                    tc_ld32(taddr + (uint32_t)(c * 32), v);
                    const uint8_t* rrow = nullptr;
                    uint8_t* srow = nullptr;
                    if (RES) {  // every warp touches each 64-channel piece exactly once (c_step == 2: no fused head with RES / TS)
                        mbar_wait(rfull_bar(c >> 1), rphase);
                        rrow = base_ptr + L::off_res + (c >> 1) * kResPiece + row * 128;
                    } else if (TS) {
                        mbar_wait(rempty_bar(c >> 1), rphase ^ 1u);  // the previous tile's store has read the piece out
                    }
                    if (TS) srow = base_ptr + L::off_res + (c >> 1) * kResPiece + row * 128;
                    if (valid)
                        epilogue_chunk<RES, TS>(p, v, n_tile * BLOCK_N + c * 32, c * 32, n, d, h, w, sc, sc + BLOCK_N, sc + 2 * BLOCK_N,
                                       sc + 3 * BLOCK_N, sc + 4 * BLOCK_N, BLOCK_N, head_acc, rrow, row & 7, PM, frame, srow);
                    if (RES || TS) {
                        if (TS) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) mbar_arrive(TS ? sdone_bar(c >> 1) : rempty_bar(c >> 1));
                    }
                }
                if (valid && p.logits) p.logits[(((long long)n * p.oD + d) * p.oH + h) * p.oW + w] = head_acc + p.head_b;
            }
            if (RES || TS) rphase ^= 1u;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(tempty_bar(acc));  // the leader's MMA thread waits for both CTAs' epilogues
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    } else if ((PM || TS) && warp == kAuxWarp) {
        // ===== auxiliary warp: polynomial coefficient tiles one pair iteration ahead (PM) and the TMA stores of finished
        // output pieces (TS) =====
        auto produce = [&](int q, int itq) {
            const int a = itq & 1;
            if (itq >= 2) mbar_wait(tfull_bar(a), (uint32_t)(((itq >> 1) - 1) & 1));  // MMAs of iteration itq - 2 are done
            int tn[2], th0[2], tw0[2];
            bool tlive[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                int mj = q * 2 + j;  // n_tiles == 1
                tlive[j] = mj < p.m_tiles;
                tw0[j] = (mj % p.tiles_w) * p.bw; mj /= p.tiles_w;
                th0[j] = (mj % p.tiles_h) * p.bh; mj /= p.tiles_h;
                tn[j] = tlive[j] ? mj / p.tiles_d : 0;
            }
            pm_produce_b<BLOCK_N / 2>(p, base_ptr + L::off_pm + 4096 + a * L::pm_b, lane, 32, rank, tn, th0, tw0, tlive);
            __syncwarp();
            if (lane == 0) mbar_arrive_leader_release(bxfull_bar(a));
        };
        if (PM) {
            for (int r = lane; r < 128; r += 32) pm_write_basis(base_ptr + L::off_pm, r, (r / p.bw) % p.bh, r % p.bw, rank);
            if (cluster_id < total_pairs) produce(cluster_id, 0);
        }
        int it = 0;
        uint32_t sphase = 0;
        for (int pt = cluster_id; pt < total_pairs; pt += n_clusters, ++it) {
            if (PM && pt + n_clusters < total_pairs) produce(pt + n_clusters, it + 1);
            if (TS) {
                const int n_tile = pt % p.n_tiles;
                int m = (pt / p.n_tiles) * 2 + (int)rank;
                const bool live = m < p.m_tiles;
                const int tw = m % p.tiles_w; m /= p.tiles_w;
                const int th = m % p.tiles_h; m /= p.tiles_h;
                const int td = m % p.tiles_d;
                const int n = m / p.tiles_d;
                for (int s = 0; s < RP; ++s) {
                    mbar_wait(sdone_bar(s), sphase);
                    if (lane == 0) {
                        if (live) {
                            const int col = n_tile * BLOCK_N + s * 64;
                            const int gq = col / p.cout, co = col - gq * p.cout;  // ConvTranspose: weight group = output offset
                            int ow = tw * p.bw * p.up, oh = th * p.bh * p.up, od = td * p.bd * p.up_d;
                            if (p.up == 2) {
                                if (p.up_d == 2) { od += gq >> 2; oh += (gq >> 1) & 1; ow += gq & 1; }
                                else { oh += gq >> 1; ow += gq & 1; }
                            }
                            tma_store_5d(&map_o, base + L::off_res + s * kResPiece, (int)p.out_coffset + co, ow, oh, od, n);
                        }
                        tma_store_commit_and_wait_read();
                        mbar_arrive(rempty_bar(s));  // free for the next residual load / the next tile's results
                    }
                    __syncwarp();
                }
                sphase ^= 1u;
            }
        }
        if (TS && lane == 0) tma_store_drain();  // stores complete before the CTA exits
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // neither CTA frees tensor memory or exits while its peer can still touch it
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------------
// conv_halo_kernel — 2D 3x3 stride-1 layers with <= 128 output channels (layer1, layer2, conv7.2, conv8.x).
//
// The per-tap kernel above re-fetches the 128-pixel A tile from L2 for each of the 9 taps; at full resolution with a
// narrow N that makes the layer L2-bandwidth bound (profiles/conv_layers_r1.md: 20-40 % tensor pipe at 13-17 TB/s of L2
// reads).  Here a CTA owns a 16x16-pixel output tile = two 16(h) x 8(w) column blocks = M 256 (two M=128 MMAs sharing
// every B tile), and the activations arrive as SLABS: for input-channel chunk c and horizontal tap dx, an
// (18 rows x 8 cols) box per column block, shifted by dx.  A slab serves the three vertical taps dy = 0,1,2 by moving
// the UMMA descriptor's start address down one pixel row (8 pixels x 128 B = 1024 B, so every 8-row swizzle group stays
// 1024-byte aligned).  L2 -> SMEM traffic per 128 output pixels drops from 9 x 16 KB (A) + 9 x B to 3 x 18 KB + 4.5 x B.
// Pipelines: A ring (one stage = the two slabs of a (chunk, dx)), B ring (one stage per tap), 2 TMEM accumulator stages
// of 2 x BLOCK_N columns.  Same warp roles, epilogue and descriptors as conv_gemm_kernel.
// ------------------------------------------------------------------------------------------------------------------
constexpr uint32_t kSlabBytes = 18 * 8 * 128;  // 18,432: (16 + 2) pixel rows x 8 pixels x 64 channels fp16

template <int BLOCK_N, int A_STAGES, int B_STAGES>
struct HaloLayout {
    static constexpr uint32_t a_stage = 2 * kSlabBytes;
    static constexpr uint32_t b_stage = BLOCK_N * 128;
    static constexpr uint32_t off_b = A_STAGES * a_stage;
    static constexpr uint32_t off_bar = off_b + B_STAGES * b_stage;
    static constexpr uint32_t off_tmem = off_bar + (2 * A_STAGES + 2 * B_STAGES + 4) * 8;
    static constexpr uint32_t off_ss = (off_tmem + 16 + 15) & ~15u;
    static constexpr uint32_t total = off_ss + 2 * 8 * BLOCK_N * 4 + 1024;
};

template <int BLOCK_N, int A_STAGES, int B_STAGES>
__global__ void __launch_bounds__(kThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2,
                 const __grid_constant__ CUtensorMap map_b, const __grid_constant__ ConvParams p) {
    using L = HaloLayout<BLOCK_N, A_STAGES, B_STAGES>;
    constexpr int TMEM_COLS = 4 * BLOCK_N;  // 2 accumulator stages x 2 column blocks
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar0 = base + L::off_bar;
    auto afull = [&](int s) { return bar0 + 8u * s; };
    auto aempty = [&](int s) { return bar0 + 8u * (A_STAGES + s); };
    auto bfull = [&](int s) { return bar0 + 8u * (2 * A_STAGES + s); };
    auto bempty = [&](int s) { return bar0 + 8u * (2 * A_STAGES + B_STAGES + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * A_STAGES + 2 * B_STAGES + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * A_STAGES + 2 * B_STAGES + 2 + s); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + L::off_tmem);
    float* ss = reinterpret_cast<float*>(base_ptr + L::off_ss);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.m_tiles;  // n_tiles == 1: BLOCK_N covers all output channels

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < A_STAGES; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
        for (int s = 0; s < B_STAGES; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + L::off_tmem),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer: same (chunk, dx, dy) order as the MMA issuer consumes =====
        if (lane == 0) {
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                int m = tile;
                const int tw = m % p.tiles_w; m /= p.tiles_w;
                const int th = m % p.tiles_h;
                const int n = m / p.tiles_h;
                const int w0 = tw * 16, h0 = th * 16;
                const int n1 = p.idx1 ? __ldg(p.idx1 + n) : n, n2 = p.idx2 ? __ldg(p.idx2 + n) : n;
                for (int c = 0; c < p.cin_chunks; ++c) {
                    const CUtensorMap* am = c < p.split_chunk ? &map_a : &map_a2;
                    const int cc = (c < p.split_chunk ? c : c - p.split_chunk) * kBlockK;
                    const int ns = c < p.split_chunk ? n1 : n2;
                    for (int dx = 0; dx < 3; ++dx) {
                        mbar_wait(aempty(sa), pa ^ 1u);
                        mbar_expect_tx(afull(sa), 2 * kSlabBytes);
                        const uint32_t dst = base + sa * L::a_stage;
                        tma_load_5d(dst, am, afull(sa), cc, w0 - 1 + dx, h0 - 1, 0, ns);
                        tma_load_5d(dst + kSlabBytes, am, afull(sa), cc, w0 + 7 + dx, h0 - 1, 0, ns);
                        if (++sa == A_STAGES) { sa = 0; pa ^= 1u; }
                        for (int dy = 0; dy < 3; ++dy) {
                            mbar_wait(bempty(sb), pb ^ 1u);
                            mbar_expect_tx(bfull(sb), L::b_stage);
                            tma_load_2d(base + L::off_b + sb * L::b_stage, &map_b, bfull(sb),
                                        ((dy * 3 + dx) * p.cin_chunks + c) * kBlockK, 0);
                            if (++sb == B_STAGES) { sb = 0; pb ^= 1u; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BLOCK_N);
            int sa = 0, sb = 0, acc = 0;
            uint32_t pa = 0, pb = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 2 * BLOCK_N);
                uint32_t first = 1;
                for (int c = 0; c < p.cin_chunks; ++c) {
                    for (int dx = 0; dx < 3; ++dx) {
                        mbar_wait(afull(sa), pa);
                        tc_fence_after();
                        const uint32_t a_base = base + sa * L::a_stage;
                        for (int dy = 0; dy < 3; ++dy) {
                            mbar_wait(bfull(sb), pb);
                            tc_fence_after();
                            const uint64_t bdesc = make_smem_desc(base + L::off_b + sb * L::b_stage);
#pragma unroll
                            for (int blk = 0; blk < 2; ++blk) {
                                // vertical tap dy = start one pixel row (8 px x 128 B) further down the slab
                                const uint64_t adesc = make_smem_desc(a_base + blk * kSlabBytes + dy * 1024);
#pragma unroll
                                for (int kk = 0; kk < kBlockK / 16; ++kk)
                                    tc_mma_f16(d_tmem + (uint32_t)(blk * BLOCK_N), adesc + (uint64_t)(kk * 2),
                                               bdesc + (uint64_t)(kk * 2), idesc, (first && kk == 0) ? 0u : 1u);
                            }
                            first = 0;
                            tc_commit(bempty(sb));
                            if (++sb == B_STAGES) { sb = 0; pb ^= 1u; }
                        }
                        tc_commit(aempty(sa));
                        if (++sa == A_STAGES) { sa = 0; pa ^= 1u; }
                    }
                }
                tc_commit(tfull_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else if (warp < kAuxWarp) {
        // ===== epilogue warps 2..9: warps 2-5 take column block 0, warps 6-9 column block 1 =====
        const int sub = warp & 3;
        const int blk = (warp - 2) >> 2;
        const int row = sub * 32 + lane;  // row of a column block: (h_l, w_l) = (row / 8, row % 8)
        const int et = threadIdx.x - 64;
        int acc = 0;
        uint32_t acc_phase = 0;
        int last_n[2] = {-1, -1};
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            int m = tile;
            const int tw = m % p.tiles_w; m /= p.tiles_w;
            const int th = m % p.tiles_h;
            const int n = m / p.tiles_h;
            float* sc = ss + acc * 8 * BLOCK_N;
            stage_tile_params<BLOCK_N>(p, sc, 0, n, et, last_n[acc] < 0, last_n[acc] != n);
            last_n[acc] = n;
            const int h = th * 16 + (row >> 3);
            const int w = tw * 16 + blk * 8 + (row & 7);
            const bool valid = h < p.H && w < p.W;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * 2 * BLOCK_N + blk * BLOCK_N);
            float head_acc = 0.f;
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                uint32_t v[32];
                tc_ld32(taddr + (uint32_t)(c * 32), v);
                if (valid)
                    epilogue_chunk<false, false>(p, v, c * 32, c * 32, n, 0, h, w, sc, sc + BLOCK_N, sc + 2 * BLOCK_N, sc + 3 * BLOCK_N,
                                   sc + 4 * BLOCK_N, BLOCK_N, head_acc);
            }
            if (valid && p.logits) p.logits[((long long)n * p.oH + h) * p.oW + w] = head_acc + p.head_b;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}


// Tile m of the pair slab kernel -> (sample n, tile td/th/tw).  Default: sample-major.  p.group > 1: samples are scheduled in
// groups of p.group consecutive samples, TILE-major inside a group, so that the tiles of the group's samples at one image
// position run at the same time: the decoder's per-map residual (res_index: the tasks of a map are consecutive samples)
// is then fetched from DRAM once per group instead of once per sample (conv8.0: 670 of 1470 MB per 10 tasks in ncu).
__device__ __forceinline__ void halo_tile(const ConvParams& p, int m, int& n, int& td, int& th, int& tw) {
    const int T = p.tiles_w * p.tiles_h * p.tiles_d;
    int t;
    if (p.group > 1) {
        const int gt = p.group * T;
        const int g = m / gt, r = m - g * gt;
        const int left = p.N - g * p.group;
        const int gs = left < p.group ? (left > 0 ? left : 1) : p.group;  // the last group may be smaller
        t = r / gs;
        n = g * p.group + (r - t * gs);
    } else {
        n = m / T;
        t = m - n * T;
    }
    tw = t % p.tiles_w; t /= p.tiles_w;
    th = t % p.tiles_h;
    td = t / p.tiles_h;
}

template <int BLOCK_N, int A_STAGES, int B_STAGES, bool RES, bool PM, bool TS, bool P3 = false>
struct HaloLayout2 {  // per CTA of a pair: its own slabs, half of every B tile
    static constexpr int ss_rows = P3 ? 16 : 8;  // parameter rows per accumulator stage (P3: + A, B of up to 6 h-classes)
    static constexpr uint32_t a_stage = 2 * kSlabBytes;
    static constexpr uint32_t b_stage = (BLOCK_N / 2) * 128;
    static constexpr uint32_t off_b = A_STAGES * a_stage;
    static constexpr uint32_t off_res = off_b + B_STAGES * b_stage;  // residual pieces [column block][64-channel chunk]
    static constexpr uint32_t off_pm = off_res + ((RES || TS) ? 2 * (BLOCK_N / 64) * kResPiece : 0);  // basis tiles (2 x 4 KB), then
    static constexpr uint32_t pm_b = (BLOCK_N / 2) * 32;                                     // 2 coefficient tiles
    static constexpr uint32_t off_bar = off_pm + (PM ? 2 * 4096 + 2 * pm_b : 0);
    static constexpr uint32_t off_tmem = off_bar + (2 * A_STAGES + 2 * B_STAGES + 4 + 3 * kMaxResPieces + 2) * 8;
    static constexpr uint32_t off_ss = (off_tmem + 16 + 15) & ~15u;
    static constexpr uint32_t total = off_ss + 2 * ss_rows * BLOCK_N * 4 + 1024;
};

// conv_halo2_kernel — the slab kernel on CTA pairs (cta_group::2): the pair owns two adjacent 16x16 tiles, every MMA is
// M256 = column block `blk` of both CTAs, and each CTA loads (and feeds from its shared memory) only half of each B tile.
// RES: TMA-staged residual tile as in conv_gemm2_kernel; a piece = (column block, 64-channel chunk) = 128 pixels x 128 B.
// PM: the interior polynomial coordinate term is one extra MMA per accumulator (pm_* above).
// TS: TMA-store epilogue (pieces shared with the residual staging; the auxiliary warp issues the stores).
// p.res_mma (= BLOCK_N / 64 extra K chunks; host sets p.res = nullptr and RES = false): the residual tile is not staged for
// the epilogue but ACCUMULATED BY THE TENSOR CORE: its two 128-pixel x 64-channel pieces per chunk are loaded into an A
// stage like slabs (same [pixel row][8 pixels][128 B] layout, vertical tap 0) and multiplied by an identity block appended
// to the weights (fp16 x 1.0 into the fp32 accumulator: exact; one N = 64 MMA per 64 output channels).  With the staged variant the residual of tile i+1 could
// only be requested once the TMA store of tile i had drained the shared piece, so consecutive epilogues were serialised
// with a store and an L2/DRAM round trip between them (ncu, conv8.0: tensor pipe 42 %, issuer waiting on the accumulator,
// epilogue waiting on the residual).  Through the A ring the residual is prefetched like any operand; K grows by 64 per
// 64 output channels.
// p.bres (host: taps x chunks == B_STAGES, 64-output-channel layers): the B ring holds ALL of the layer's weight tiles, one
// per slot, so the producer fills it during the CTA's first tile and the issuers never wait for or release a B stage
// again.  These layers are bound by the operand stream from L2 (profiles/conv_decoder_r2.md: conv8.2 pulls 216 KB of
// slabs + 72 KB of weights per tile at 9 TB/s aggregate); the weights were a quarter of it.
// p.fold: ConvTranspose(k=2, s=2) folded into the 3x3 conv that consumes it (train.py:196-201: upconv7 -> conv7.0,
// upconv8 -> conv8.0; no nonlinearity in between, so the composition is linear).  Output pixel (2i + a, 2j + b) of the conv
// sees the upsampled pixels (2i + a + dy, 2j + b + dx), which come from the 2 x 2 low-resolution pixels (i + a - 1 .. i + a,
// j + b - 1 .. j + b): per output parity (a, b) the pair of layers is ONE 2x2-tap convolution over the low-resolution
// input with composite weights W'[par][tap] = sum of Wconv[dy, dx] * Wup[sub-position] (host: cnn.py fold_weights).
// The tile space is the low-resolution grid; a pair iteration is (tile pair q, N half nh, parity par); slabs start at
// (h0 - 1 + a, w0 - 1 + b + dx) and serve two vertical taps; residual tiles are read and results stored through tensor
// maps with element stride 2 starting at (2 h0 + a, 2 w0 + b).  The ConvTranspose bias reaches the output through the
// taps that fall inside the image, a per-border-case constant that the host adds to the polynomial's constant term
// (the polynomial itself is re-expressed per parity in tile-space coordinates: nrrt_poly_fold).  20 % fewer MACs than
// the two layers, and the upsampled activation (the widest tensor of the stage) is never written or read.
template <int BLOCK_N, int A_STAGES, int B_STAGES, bool RES, bool PM, bool TS, bool P3 = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_halo2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2,
                 const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_r,
                 const __grid_constant__ CUtensorMap map_o, const __grid_constant__ ConvParams p) {
    using L = HaloLayout2<BLOCK_N, A_STAGES, B_STAGES, RES, PM, TS, P3>;
    constexpr int RC = BLOCK_N / 64;  // 64-channel chunks per column block; residual pieces = 2 * RC
    constexpr int TMEM_COLS = 4 * BLOCK_N;  // 2 accumulator stages x 2 column blocks
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar0 = base + L::off_bar;
    auto afull = [&](int s) { return bar0 + 8u * s; };
    auto aempty = [&](int s) { return bar0 + 8u * (A_STAGES + s); };
    auto bfull = [&](int s) { return bar0 + 8u * (2 * A_STAGES + s); };
    auto bempty = [&](int s) { return bar0 + 8u * (2 * A_STAGES + B_STAGES + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * A_STAGES + 2 * B_STAGES + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * A_STAGES + 2 * B_STAGES + 2 + s); };
    auto rfull_bar = [&](int s) { return bar0 + 8u * (2 * A_STAGES + 2 * B_STAGES + 4 + s); };
    auto rempty_bar = [&](int s) { return bar0 + 8u * (2 * A_STAGES + 2 * B_STAGES + 4 + kMaxResPieces + s); };
    auto sdone_bar = [&](int s) { return bar0 + 8u * (2 * A_STAGES + 2 * B_STAGES + 4 + 2 * kMaxResPieces + s); };
    auto bxfull_bar = [&](int s) { return bar0 + 8u * (2 * A_STAGES + 2 * B_STAGES + 4 + 3 * kMaxResPieces + s); };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + L::off_tmem);
    float* ss = reinterpret_cast<float*>(base_ptr + L::off_ss);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the pair's MMAs), 1 = peer
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    // a CTA pair owns two consecutive 16x16 tiles; n_tiles == 1 unless folded: pair iteration = (tile pair, N half, parity)
    const int total_pairs = ((p.m_tiles + 1) >> 1) * (p.fold ? 4 * p.n_tiles : 1);
    auto decode = [&](int pt, int& q, int& nh, int& par) {
        if (p.fold) { par = pt & 3; pt >>= 2; nh = pt % p.n_tiles; q = pt / p.n_tiles; }
        else { par = 0; nh = 0; q = pt; }
    };

    if (warp == 1 && lane == 0) {
        // "empty" and accumulator-full barriers collect the commits of both MMA issuers
        for (int s = 0; s < A_STAGES; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 2); }
        for (int s = 0; s < B_STAGES; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 2); }
        for (int s = 0; s < 2; ++s) { mbar_init(tfull_bar(s), 2); mbar_init(tempty_bar(s), 2 * kEpiWarps); }
        if (RES || TS)  // piece free: the four warps of its column block, or (TS) the store warp once the piece has been read out
            for (int s = 0; s < 2 * RC; ++s) {
                mbar_init(rfull_bar(s), 1); mbar_init(rempty_bar(s), TS ? 1 : kEpiWarps / 2); mbar_init(sdone_bar(s), kEpiWarps / 2);
            }
        if (PM) for (int s = 0; s < 2; ++s) mbar_init(bxfull_bar(s), 2);  // one arrival per CTA of the pair
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(base + L::off_tmem),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer: same (chunk, dx, dy) order as the MMA issuer consumes =====
        if (lane == 0) {
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0, rphase = 0;
            for (int pt = cluster_id; pt < total_pairs; pt += n_clusters) {
                int q, nh, par;
                decode(pt, q, nh, par);
                const int fa = par >> 1, fb = par & 1;  // output parity (row, column) of a folded layer
                const int m = q * 2 + (int)rank;
                const bool live = m < p.m_tiles;
                int n = 0, td = 0, th = 0, tw = 0;  // 3D: one depth slice per tile (tiles_d = D)
                if (live) halo_tile(p, m, n, td, th, tw);
                const int w0 = tw * 16, h0 = th * 16;
                const int n1 = !live ? 0x3fffffff : (p.idx1 ? __ldg(p.idx1 + n) : n);
                const int n2 = !live ? 0x3fffffff : (p.idx2 ? __ldg(p.idx2 + n) : n);
                const int nr = ((!RES && !p.res_mma) || !live) ? 0x3fffffff : (p.res_idx ? __ldg(p.res_idx + n) : n);
                if ((RES || p.res_mma) && live)  // the residual pieces are loaded behind this tile's operands: start their trip to L2 now
                    for (int s = 0; s < 2 * RC; ++s) {
                        const int rw = w0 + (s / RC) * 8;
                        tma_prefetch_5d(&map_r, (int)p.res_coffset + nh * BLOCK_N + (s % RC) * 64, p.fold ? 2 * rw + fb : rw,
                                        p.fold ? 2 * h0 + fa : h0, td, nr);
                    }
                for (int c = 0; c < p.cin_chunks; ++c) {
                    const CUtensorMap* am = c < p.split_chunk ? &map_a : &map_a2;
                    const int cc = (c < p.split_chunk ? c : c - p.split_chunk) * kBlockK;
                    const int ns = c < p.split_chunk ? n1 : n2;
                    for (int o = 0; o < p.outer; ++o) {  // (dz, dx): 3 slabs per chunk in 2D, 9 in 3D (depth slice d + dz - 1)
                        const int dx = p.fold ? o : o % 3, dz = p.fold ? 0 : o / 3;
                        const int dd = td + dz - (p.outer == 9 ? 1 : 0);
                        const int xs = w0 - 1 + dx + fb, ys = h0 - 1 + fa;
                        mbar_wait(aempty(sa), pa ^ 1u);
                        if (rank == 0) mbar_expect_tx(afull(sa), 4 * kSlabBytes);  // two slabs from each CTA of the pair
                        const uint32_t dst = base + sa * L::a_stage;
                        const uint32_t abar = afull(sa) & kPeerBitMask;
                        tma2_load_5d(dst, am, abar, cc, xs, ys, dd, ns);
                        tma2_load_5d(dst + kSlabBytes, am, abar, cc, xs + 8, ys, dd, ns);
                        if (++sa == A_STAGES) { sa = 0; pa ^= 1u; }
                        for (int dy = 0; dy < p.inner; ++dy) {
                            if (!p.bres || pt == cluster_id) {  // resident weights: only the CTA's first tile loads them
                                mbar_wait(bempty(sb), pb ^ 1u);
                                if (rank == 0) mbar_expect_tx(bfull(sb), 2 * L::b_stage);  // each CTA loads half of the B tile
                                const int tap = p.fold ? (par * 4 + dy * 2 + dx) : ((dz * 3 + dy) * 3 + dx);
                                tma2_load_2d(base + L::off_b + sb * L::b_stage, &map_b, bfull(sb) & kPeerBitMask,
                                             (tap * p.cin_chunks + c) * kBlockK, nh * BLOCK_N + (int)rank * (BLOCK_N / 2));
                            }
                            if (++sb == B_STAGES) { sb = 0; pb ^= 1u; }
                        }
                    }
                }
                for (int kc = 0; kc < p.res_mma; ++kc) {  // residual chunk kc as one more A stage + its identity B tile
                    mbar_wait(aempty(sa), pa ^ 1u);
                    if (rank == 0) mbar_expect_tx(afull(sa), 4 * kResPiece);  // two pieces from each CTA of the pair
                    const uint32_t dst = base + sa * L::a_stage;
                    const uint32_t abar = afull(sa) & kPeerBitMask;
                    const int rc0 = (int)p.res_coffset + nh * BLOCK_N + kc * 64;
                    const int rx = p.fold ? 2 * w0 + fb : w0, ry = p.fold ? 2 * h0 + fa : h0, rstep = p.fold ? 16 : 8;
                    tma2_load_5d(dst, &map_r, abar, rc0, rx, ry, td, nr);
                    tma2_load_5d(dst + kSlabBytes, &map_r, abar, rc0, rx + rstep, ry, td, nr);
                    if (++sa == A_STAGES) { sa = 0; pa ^= 1u; }
                    mbar_wait(bempty(sb), pb ^ 1u);
                    if (rank == 0) mbar_expect_tx(bfull(sb), 2 * L::b_stage);
                    // identity rows of output channels [kc * 64, +64): an N = 64 MMA, each CTA feeding 32 of them (the tile's
                    // other rows are loaded but not read)
                    tma2_load_2d(base + L::off_b + sb * L::b_stage, &map_b, bfull(sb) & kPeerBitMask,
                                 p.res_k0 + kc * kBlockK, nh * BLOCK_N + kc * 64 + (int)rank * 32);
                    if (++sb == B_STAGES) { sb = 0; pb ^= 1u; }
                }
                if (RES) {
                    for (int s = 0; s < 2 * RC; ++s) {
                        mbar_wait(rempty_bar(s), rphase ^ 1u);
                        mbar_expect_tx(rfull_bar(s), kResPiece);
                        const int rw = w0 + (s / RC) * 8;
                        tma_load_5d(base + L::off_res + s * kResPiece, &map_r, rfull_bar(s),
                                    (int)p.res_coffset + nh * BLOCK_N + (s % RC) * 64, p.fold ? 2 * rw + fb : rw,
                                    p.fold ? 2 * h0 + fa : h0, td, nr);
                    }
                    rphase ^= 1u;
                }
            }
            // drain both rings: multicast commits of the leader must not land in a CTA that already exited
            for (int i = 0; i < A_STAGES; ++i) { mbar_wait(aempty(sa), pa ^ 1u); if (++sa == A_STAGES) { sa = 0; pa ^= 1u; } }
            if (!p.bres)  // (resident weights are never released: nothing of the leader's lands on these barriers)
                for (int i = 0; i < B_STAGES; ++i) { mbar_wait(bempty(sb), pb ^ 1u); if (++sb == B_STAGES) { sb = 0; pb ^= 1u; } }
        }
    } else if (warp == 1 || warp == kIssue2Warp) {
        // ===== two MMA issuers, one per column block.  A single thread needs ~100 cycles of dependent instructions per
        // tcgen05.mma (descriptor moves into uniform registers + the elect wrapper), which capped every N <= 128 layer at the
        // issue rate instead of the tensor pipe; the two column blocks accumulate into disjoint tensor-memory columns, so
        // they are issued independently.  Every smem-stage / accumulator barrier therefore counts two commits. =====
        if (rank == 0) {  // the whole warp runs the loop (uniform control flow); one elected lane issues
            constexpr uint32_t idesc = make_idesc(BLOCK_N, 256);  // M 256 = column block `blk` of both CTAs
            const int blk = warp == 1 ? 0 : 1;
            int sa = 0, sb = 0, acc = 0;
            uint32_t pa = 0, pb = 0, acc_phase = 0;
            for (int pt = cluster_id; pt < total_pairs; pt += n_clusters) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 2 * BLOCK_N + blk * BLOCK_N);
                uint32_t first = 1;
                for (int c = 0; c < p.cin_chunks; ++c) {
                    for (int o = 0; o < p.outer; ++o) {
                        mbar_wait(afull(sa), pa);
                        tc_fence_after();
                        const uint32_t a_base = base + sa * L::a_stage + blk * kSlabBytes;
                        for (int dy = 0; dy < p.inner; ++dy) {
                            if (!p.bres || pt == cluster_id) {  // resident weights arrive once, during the first tile
                                mbar_wait(bfull(sb), pb);
                                tc_fence_after();
                            }
                            const uint64_t bdesc = make_smem_desc(base + L::off_b + sb * L::b_stage);
                            // vertical tap dy = start one pixel row (8 px x 128 B) further down the slab
                            const uint64_t adesc = make_smem_desc(a_base + dy * 1024);
                            if (elect_one()) {
#pragma unroll
                                for (int kk = 0; kk < kBlockK / 16; ++kk)
                                    tc_mma2_f16(d_tmem, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc,
                                               (first && kk == 0) ? 0u : 1u);
                                if (!p.bres) tc_commit2(bempty(sb));
                            }
                            __syncwarp();
                            first = 0;
                            if (++sb == B_STAGES) { sb = 0; pb ^= 1u; }
                        }
                        if (elect_one()) tc_commit2(aempty(sa));
                        __syncwarp();
                        if (++sa == A_STAGES) { sa = 0; pa ^= 1u; }
                    }
                }
                for (int kc = 0; kc < p.res_mma; ++kc) {  // + residual chunk x identity
                    mbar_wait(afull(sa), pa);
                    mbar_wait(bfull(sb), pb);
                    tc_fence_after();
                    const uint64_t bdesc = make_smem_desc(base + L::off_b + sb * L::b_stage);
                    const uint64_t adesc = make_smem_desc(base + sa * L::a_stage + blk * kSlabBytes);
                    if (elect_one()) {  // D[:, kc * 64 : +64] += R_kc x I (first == 0 here: the conv taps came before)
                        constexpr uint32_t idesc64 = make_idesc(64, 256);
#pragma unroll
                        for (int kk = 0; kk < kBlockK / 16; ++kk)
                            tc_mma2_f16(d_tmem + (uint32_t)(kc * 64), adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), idesc64, 1u);
                        tc_commit2(bempty(sb));
                        tc_commit2(aempty(sa));
                    }
                    __syncwarp();
                    first = 0;
                    if (++sb == B_STAGES) { sb = 0; pb ^= 1u; }
                    if (++sa == A_STAGES) { sa = 0; pa ^= 1u; }
                }
                if (PM) {
                    mbar_wait_cluster(bxfull_bar(acc), acc_phase);  // both CTAs' coefficient tiles of this pair iteration
                    tc_fence_after();
                    if (elect_one())
                        tc_mma2_f16(d_tmem, make_smem_desc_nosw(base + L::off_pm + blk * 4096, 2048, 128),
                                   make_smem_desc_nosw(base + L::off_pm + 2 * 4096 + acc * L::pm_b, (BLOCK_N / 2) * 16, 128), idesc, 1u);
                    __syncwarp();
                }
                if (elect_one()) tc_commit2(tfull_bar(acc));
                __syncwarp();
                if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
            }
        }
    } else if (warp < kAuxWarp) {
        // ===== epilogue warps 2..9: warps 2-5 take column block 0, warps 6-9 column block 1 =====
        const int sub = warp & 3;
        const int blk = (warp - 2) >> 2;
        const int row = sub * 32 + lane;  // row of a column block: (h_l, w_l) = (row / 8, row % 8)
        const int et = threadIdx.x - 64;
        int acc = 0;
        uint32_t acc_phase = 0, rphase = 0;
        int last_n[2] = {-1, -1}, last_nh[2] = {-1, -1};
        for (int pt = cluster_id; pt < total_pairs; pt += n_clusters) {
            int q, nh, par;
            decode(pt, q, nh, par);
            const int m = q * 2 + (int)rank;
            const bool live = m < p.m_tiles;
            int n = 0, td = 0, th = 0, tw = 0;
            if (live) halo_tile(p, m, n, td, th, tw);
            // folded layers: the polynomial is stored per (sample, parity); n only indexes it (results leave by TMA)
            if (p.fold) n = n * 4 + par;
            const int col0 = nh * BLOCK_N;
            float* sc = ss + acc * L::ss_rows * BLOCK_N;
            stage_tile_params<BLOCK_N>(p, sc, col0, n, et, last_nh[acc] != nh, !PM && (last_n[acc] != n || last_nh[acc] != nh));
            last_n[acc] = n; last_nh[acc] = nh;
            if (P3 && live) stage_poly3<BLOCK_N>(p, sc, n, td, th * 16, et, col0);
            int frame = 0;
            if (PM) {
                const int h0 = th * 16, w0 = tw * 16;
                const int hr = h0 == 0 ? 0 : (p.H - 1 - h0 < 16 ? p.H - 1 - h0 : -1);
                const int wc = w0 == 0 ? 0 : (p.W - 1 - w0 < 16 ? p.W - 1 - w0 : -1);
                if (live && (hr >= 0 || wc >= 0)) {
                    pm_stage_frame<BLOCK_N>(p, sc, n, et, h0, w0, hr, wc, col0);
                    const bool on_r = (row >> 3) == hr, on_c = blk * 8 + (row & 7) == wc;
                    frame = on_r ? (on_c ? 3 : 1) : (on_c ? 2 : 0);
                }
            }

            const int h = th * 16 + (row >> 3);
            const int w = tw * 16 + blk * 8 + (row & 7);
            const bool valid = live && h < p.H && w < p.W;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(acc * 2 * BLOCK_N + blk * BLOCK_N);
            float head_acc = 0.f;
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / 32; ++c) {
                uint32_t v[32];
                tc_ld32(taddr + (uint32_t)(c * 32), v);
                const int piece = blk * RC + (c >> 1);
                const uint8_t* rrow = nullptr;
                uint8_t* srow = nullptr;
                if (RES) {
                    if (!(c & 1)) mbar_wait(rfull_bar(piece), rphase);
                    rrow = base_ptr + L::off_res + piece * kResPiece + row * 128;
                } else if (TS) {
                    if (!(c & 1)) mbar_wait(rempty_bar(piece), rphase ^ 1u);  // the previous tile's store has read the piece out
                }
                if (TS) srow = base_ptr + L::off_res + piece * kResPiece + row * 128;
                if (valid && !(p.dbg & 2))
                    epilogue_chunk<RES, TS>(p, v, col0 + c * 32, c * 32, n, td, h, w, sc, sc + BLOCK_N, sc + 2 * BLOCK_N, sc + 3 * BLOCK_N,
                                   sc + 4 * BLOCK_N, BLOCK_N, head_acc, rrow, row & 7, PM, frame, srow,
                                   P3 ? hclass3(h, p.H) - hclass3(th * 16, p.H) : 0);
                if ((RES || TS) && (c & 1)) {  // both 32-column halves of the piece are done
                    if (TS) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(TS ? sdone_bar(piece) : rempty_bar(piece));
                }
            }
            if (RES || TS) rphase ^= 1u;
            if (valid && p.logits) p.logits[(((long long)n * p.oD + td) * p.oH + h) * p.oW + w] = head_acc + p.head_b;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
    } else if ((PM || TS) && warp == kAuxWarp) {
        // ===== auxiliary warp: polynomial coefficient tiles one pair iteration ahead (PM) and the TMA stores of finished
        // output pieces (TS) =====
        auto produce = [&](int q, int itq) {  // coefficient tile of pair iteration q (the itq-th of this pair) into stage itq & 1
            const int a = itq & 1;
            // the stage was last read by the MMAs of iteration itq - 2: wait for that accumulator's commit
            if (itq >= 2) mbar_wait(tfull_bar(a), (uint32_t)(((itq >> 1) - 1) & 1));
            int tn[2], th0[2], tw0[2];
            bool tlive[2];
            int qq, nh, par;
            decode(q, qq, nh, par);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int mj = qq * 2 + j;
                tlive[j] = mj < p.m_tiles;
                int nj = 0, tdj = 0, thj = 0, twj = 0;
                if (tlive[j]) halo_tile(p, mj, nj, tdj, thj, twj);
                tw0[j] = twj * 16;
                th0[j] = thj * 16;
                tn[j] = p.fold ? nj * 4 + par : nj;
            }
            pm_produce_b<BLOCK_N / 2>(p, base_ptr + L::off_pm + 2 * 4096 + a * L::pm_b, lane, 32, rank, tn, th0, tw0, tlive,
                                      nh * BLOCK_N);
            __syncwarp();
            if (lane == 0) mbar_arrive_leader_release(bxfull_bar(a));  // one arrival per CTA of the pair
        };
        if (PM) {
            for (int r = lane; r < 256; r += 32)
                pm_write_basis(base_ptr + L::off_pm + (r >> 7) * 4096, r & 127, (r & 127) >> 3, (r >> 7) * 8 + (r & 7), rank);
            if (cluster_id < total_pairs) produce(cluster_id, 0);
        }
        int it = 0, prev_piece = -1;
        uint32_t sphase = 0;
        for (int pt = cluster_id; pt < total_pairs; pt += n_clusters, ++it) {
            if (PM && pt + n_clusters < total_pairs) produce(pt + n_clusters, it + 1);
            if (TS) {
                int q, nh, par;
                decode(pt, q, nh, par);
                const int m = q * 2 + (int)rank;
                const bool live = m < p.m_tiles;
                int n = 0, td = 0, th = 0, tw = 0;
                if (live) halo_tile(p, m, n, td, th, tw);
                for (int k = 0; k < RC; ++k)
                    for (int b = 0; b < 2; ++b) {  // order in which the epilogue finishes them
                        const int piece = b * RC + k;
                        mbar_wait(sdone_bar(piece), sphase);
                        if (lane == 0) {
                            if (live && !(p.dbg & 1)) {
                                const int ow = tw * 16 + b * 8, oh = th * 16;
                                tma_store_5d(&map_o, base + L::off_res + piece * kResPiece, (int)p.out_coffset + nh * BLOCK_N + k * 64,
                                             p.fold ? 2 * ow + (par & 1) : ow, p.fold ? 2 * oh + (par >> 1) : oh, td, n);
                            }
                            if (RES) {  // the next residual load of this piece waits for it: free it as early as possible
                                tma_store_commit_and_wait_read();
                                mbar_arrive(rempty_bar(piece));
                            } else {    // only the next tile's results need the piece: keep one store in flight
                                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                                if (prev_piece >= 0) {
                                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                                    mbar_arrive(rempty_bar(prev_piece));
                                }
                                prev_piece = piece;
                            }
                        }
                        __syncwarp();
                    }
                sphase ^= 1u;
            }
        }
        if (TS && lane == 0) tma_store_drain();  // stores complete before the CTA exits
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
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

// Every launcher has the same signature so that a prepared layer can keep a pointer to the one it resolved to.  The
// dynamic-shared-memory attribute of an instantiation is set once per device, not on every launch.
typedef int (*LaunchFn)(const CUtensorMap& ma, const CUtensorMap& ma2, const CUtensorMap& mb, const CUtensorMap& mr,
                        const CUtensorMap& mo, const ConvParams& p, int dev, int sms, cudaStream_t st);

template <typename K>
int set_smem_once(K kern, int bytes, int dev, int& done_dev) {
    if (done_dev != dev) {
        NRRT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        done_dev = dev;
    }
    return NRRT_OK;
}

template <int BLOCK_N, int STAGES>
int launch_conv(const CUtensorMap& ma, const CUtensorMap& ma2, const CUtensorMap& mb, const CUtensorMap&, const CUtensorMap&,
                const ConvParams& p, int dev, int sms, cudaStream_t st) {
    using L = SmemLayout<BLOCK_N, STAGES>;
    auto kern = conv_gemm_kernel<BLOCK_N, STAGES>;
    static int done_dev = -1;
    if (int rc = set_smem_once(kern, (int)L::total, dev, done_dev)) return rc;
    const int total = p.m_tiles * p.n_tiles;
    const int grid = total < sms ? total : sms;
    kern<<<grid, kThreads, L::total, st>>>(ma, ma2, mb, p);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

template <int BLOCK_N, int A_STAGES, int B_STAGES, bool RES, bool PM = false, bool TS = false, bool P3 = false>
int launch_halo2(const CUtensorMap& ma, const CUtensorMap& ma2, const CUtensorMap& mb, const CUtensorMap& mr, const CUtensorMap& mo,
                 const ConvParams& p, int dev, int sms, cudaStream_t st) {
    using L = HaloLayout2<BLOCK_N, A_STAGES, B_STAGES, RES, PM, TS, P3>;
    static_assert(L::total <= 232448, "shared memory budget");
    auto kern = conv_halo2_kernel<BLOCK_N, A_STAGES, B_STAGES, RES, PM, TS, P3>;
    static int done_dev = -1;
    if (int rc = set_smem_once(kern, (int)L::total, dev, done_dev)) return rc;
    const int pairs = ((p.m_tiles + 1) / 2) * (p.fold ? 4 * p.n_tiles : 1);
    const int grid = 2 * (pairs < sms / 2 ? pairs : sms / 2);
    kern<<<grid, kThreads, L::total, st>>>(ma, ma2, mb, mr, mo, p);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

template <int BLOCK_N, int STAGES, bool RES, bool PM = false, bool TS = false>
int launch_conv2(const CUtensorMap& ma, const CUtensorMap& ma2, const CUtensorMap& mb, const CUtensorMap& mr, const CUtensorMap& mo,
                 const ConvParams& p, int dev, int sms, cudaStream_t st) {
    using L = SmemLayout2<BLOCK_N, STAGES, RES, PM, TS>;
    static_assert(L::total <= 232448, "shared memory budget");
    auto kern = conv_gemm2_kernel<BLOCK_N, STAGES, RES, PM, TS>;
    static int done_dev = -1;
    if (int rc = set_smem_once(kern, (int)L::total, dev, done_dev)) return rc;
    const int pairs = ((p.m_tiles + 1) / 2) * p.n_tiles;
    const int grid = 2 * (pairs < sms / 2 ? pairs : sms / 2);  // whole CTA pairs (static cluster dims 2x1x1)
    kern<<<grid, kThreads, L::total, st>>>(ma, ma2, mb, mr, mo, p);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

template <int BLOCK_N, int A_STAGES, int B_STAGES>
int launch_halo(const CUtensorMap& ma, const CUtensorMap& ma2, const CUtensorMap& mb, const CUtensorMap&, const CUtensorMap&,
                const ConvParams& p, int dev, int sms, cudaStream_t st) {
    using L = HaloLayout<BLOCK_N, A_STAGES, B_STAGES>;
    auto kern = conv_halo_kernel<BLOCK_N, A_STAGES, B_STAGES>;
    static int done_dev = -1;
    if (int rc = set_smem_once(kern, (int)L::total, dev, done_dev)) return rc;
    const int grid = p.m_tiles < sms ? p.m_tiles : sms;
    kern<<<grid, kThreads, L::total, st>>>(ma, ma2, mb, p);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

void pick_box(int W, int H, int D, int& bw, int& bh, int& bd) {
    // Largest box of <= 128 pixels that tiles the grid with the least waste (dims need not be powers of two).
    double best = -1;
    bw = bh = bd = 1;
    for (int d = 1; d <= std::min(D, 128); ++d)
        for (int h = 1; h <= std::min(H, 128 / d); ++h) {
            const int w = std::min(W, 128 / (d * h));
            if (w < 1) continue;
            const double tiles = (double)((W + w - 1) / w) * ((H + h - 1) / h) * ((D + d - 1) / d);
            const double eff = (double)W * H * D / (tiles * 128.0);
            // among equally full tilings prefer the smallest halo (taps of neighbouring rows re-read the same pixels)
            const double halo = (double)(w + 2) * (h + 2) * (D > 1 ? d + 2 : 1) / ((double)w * h * d);
            const double score = eff - 1e-3 * halo;
            if (score > best) { best = score; bw = w; bh = h; bd = d; }
        }
}

// A layer description resolved once: tensor maps, kernel parameters and the kernel instance.  Encoding 3-5 CUtensorMaps per
// launch (a driver call each) for ~1760 launches per 4096-task step was ~35 ms of host time per step on the launching
// thread; the buffers of a forward are stable (cnn.py caches them), so the same descriptions recur every micro-batch.
struct PreparedConv {
    CUtensorMap ma, ma2, mb, mr, mo;
    ConvParams p;
    LaunchFn launch;
    double flops, bytes;  // algorithmic: 2 x MACs over the channels read; every operand and the output once
};

// ---- launch profiler (nrrt_conv_profile_*): counts every launch, brackets every n-th with CUDA events ----
struct ProfRecord { cudaEvent_t a, b; double flops, bytes; int tag; };
struct ConvProfile {
    bool on = false;
    int every = 1;
    long long seen = 0;
    double flops = 0.0, bytes = 0.0;
    std::vector<ProfRecord> recs;
    std::vector<cudaEvent_t> pool;  // events of earlier sessions, reused
};
ConvProfile& profile() {
    static thread_local ConvProfile p;
    return p;
}

int prepare_conv(const NrrtConvLayer* l, PreparedConv& out);

// The per-call pointers that never enter a tensor map or a kernel choice beyond being null or not: the cache key holds
// only their null-ness, and a cache hit patches their current values into the prepared parameters.
void patch_volatile(const NrrtConvLayer* l, ConvParams& p) {
    p.idx1 = l->in_index; p.idx2 = l->in2_index;
    if (p.res_idx) p.res_idx = l->res_index;
    if (p.poly) p.poly = l->poly;
    if (p.poly3) p.poly3 = l->poly3;
    p.logits = l->logits;
    p.head_w = l->head_weight; p.head_b = l->head_bias;
    p.group = l->res_group > 1 ? l->res_group : 1;
}

struct ConvKey {
    NrrtConvLayer l;
    int dev;
    bool operator==(const ConvKey& o) const { return dev == o.dev && memcmp(&l, &o.l, sizeof(l)) == 0; }
};
struct ConvKeyHash {
    size_t operator()(const ConvKey& k) const {
        const unsigned char* b = reinterpret_cast<const unsigned char*>(&k.l);
        uint64_t h = 1469598103934665603ull ^ (uint64_t)k.dev;
        for (size_t i = 0; i < sizeof(k.l); ++i) h = (h ^ b[i]) * 1099511628211ull;
        return (size_t)h;
    }
};

}  // namespace

extern "C" int nrrt_conv_layer(const NrrtConvLayer* l, void* stream) {
    NRRT_REQUIRE(l, "null layer");
    if (l->poly) NRRT_REQUIRE(((uintptr_t)l->poly & 15) == 0, "poly must be 16-byte aligned");
    if (l->poly3) NRRT_REQUIRE(((uintptr_t)l->poly3 & 15) == 0, "poly3 must be 16-byte aligned");
    int dev = 0;
    NRRT_CUDA_TRY(cudaGetDevice(&dev));
    static thread_local std::unordered_map<ConvKey, PreparedConv, ConvKeyHash> cache;
    static thread_local int sms_of[64] = {0};
    ConvKey key;
    memset(&key, 0, sizeof(key));
    key.l = *l;
    key.dev = dev;
    auto flag = [](const void* q) { return q ? reinterpret_cast<void*>(1) : nullptr; };
    key.l.in_index = (const int32_t*)flag(l->in_index); key.l.in2_index = (const int32_t*)flag(l->in2_index);
    key.l.res_index = (const int32_t*)flag(l->res_index);
    key.l.poly = (const float*)flag(l->poly); key.l.poly3 = (const float*)flag(l->poly3);
    key.l.logits = (float*)flag(l->logits);
    key.l.head_weight = (const float*)flag(l->head_weight); key.l.head_bias = 0.f;
    key.l.res_group = l->res_group > 1 ? 2 : 0;
    auto it = cache.find(key);
    if (it == cache.end()) {
        PreparedConv prep;
        if (int rc = prepare_conv(l, prep)) return rc;
        if (cache.size() >= 4096) cache.clear();  // descriptions of buffers that no longer exist
        it = cache.emplace(key, prep).first;
    }
    PreparedConv& pc = it->second;
    patch_volatile(l, pc.p);
    int& sms = sms_of[dev & 63];
    if (!sms) NRRT_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    ConvProfile& prof = profile();
    if (!prof.on) return pc.launch(pc.ma, pc.ma2, pc.mb, pc.mr, pc.mo, pc.p, dev, sms, st);
    ++prof.seen;
    prof.flops += pc.flops;
    prof.bytes += pc.bytes;
    if (prof.seen % prof.every != 0) return pc.launch(pc.ma, pc.ma2, pc.mb, pc.mr, pc.mo, pc.p, dev, sms, st);
    ProfRecord r;
    r.flops = pc.flops; r.bytes = pc.bytes; r.tag = l->tag;
    for (cudaEvent_t* e : {&r.a, &r.b}) {
        if (!prof.pool.empty()) { *e = prof.pool.back(); prof.pool.pop_back(); }
        else NRRT_CUDA_TRY(cudaEventCreate(e));
    }
    NRRT_CUDA_TRY(cudaEventRecord(r.a, st));
    const int rc = pc.launch(pc.ma, pc.ma2, pc.mb, pc.mr, pc.mo, pc.p, dev, sms, st);
    NRRT_CUDA_TRY(cudaEventRecord(r.b, st));
    prof.recs.push_back(r);
    return rc;
}

extern "C" int nrrt_conv_profile_start(int sample_every) {
    ConvProfile& prof = profile();
    for (const ProfRecord& r : prof.recs) { prof.pool.push_back(r.a); prof.pool.push_back(r.b); }
    prof.recs.clear();
    prof.on = true;
    prof.every = sample_every > 1 ? sample_every : 1;
    prof.seen = 0;
    prof.flops = prof.bytes = 0.0;
    return NRRT_OK;
}

extern "C" int nrrt_conv_profile_stop(void) {
    profile().on = false;
    return NRRT_OK;
}

extern "C" int nrrt_conv_profile_resume(void) {
    profile().on = true;
    return NRRT_OK;
}

extern "C" int nrrt_conv_profile_read(int64_t* launches_seen, double* flops_seen, double* bytes_seen, int max_records,
                                      int32_t* tags, float* ms, double* flops, double* bytes, int32_t* n_records) {
    ConvProfile& prof = profile();
    if (launches_seen) *launches_seen = prof.seen;
    if (flops_seen) *flops_seen = prof.flops;
    if (bytes_seen) *bytes_seen = prof.bytes;
    const int n = (int)prof.recs.size() < max_records ? (int)prof.recs.size() : (max_records > 0 ? max_records : 0);
    for (int i = 0; i < n; ++i) {
        const ProfRecord& r = prof.recs[i];
        float t = 0.f;
        NRRT_CUDA_TRY(cudaEventElapsedTime(&t, r.a, r.b));  // cudaErrorNotReady until the stream has passed the events
        if (tags) tags[i] = r.tag;
        if (ms) ms[i] = t;
        if (flops) flops[i] = r.flops;
        if (bytes) bytes[i] = r.bytes;
    }
    if (n_records) *n_records = (int)prof.recs.size();
    return NRRT_OK;
}

namespace {

int prepare_conv(const NrrtConvLayer* l, PreparedConv& out) {
    NRRT_REQUIRE(l && l->in && l->weight, "null pointer in conv layer");
    NRRT_REQUIRE((l->scale == nullptr) == (l->shift == nullptr), "scale and shift go together (both NULL = identity)");
    NRRT_REQUIRE(l->scale || !(l->post_scale || l->logits), "the identity affine is for plain stored outputs");
    NRRT_REQUIRE(l->dims == 2 || l->dims == 3, "dims must be 2 or 3");
    NRRT_REQUIRE(l->kind >= 0 && l->kind <= 5, "unknown conv kind %d", l->kind);
    NRRT_REQUIRE(l->cin > 0 && l->cin % 64 == 0, "Cin must be a multiple of 64 (got %d)", l->cin);
    NRRT_REQUIRE(l->cin2 >= 0 && l->cin2 % 64 == 0, "Cin2 must be a multiple of 64 (got %d)", l->cin2);
    if (l->cin2 > 0) NRRT_REQUIRE(l->in2 && l->in2_cstride % 8 == 0 && ((uintptr_t)l->in2 & 15) == 0 && l->in2_n >= 1, "bad second input");
    NRRT_REQUIRE(l->in_n >= 0, "bad in_n");
    NRRT_REQUIRE(l->cout > 0 && l->cout % 32 == 0, "Cout must be a multiple of 32 (got %d)", l->cout);
    NRRT_REQUIRE(l->in_cstride % 8 == 0 && l->out_cstride % 8 == 0 && l->out_coffset % 8 == 0, "channel strides/offsets must be multiples of 8");
    NRRT_REQUIRE(((uintptr_t)l->in & 15) == 0 && ((uintptr_t)l->weight & 15) == 0 && ((uintptr_t)l->out & 15) == 0, "pointers must be 16-byte aligned");
    if (l->res) NRRT_REQUIRE(l->res_cstride % 8 == 0 && l->res_coffset % 8 == 0 && ((uintptr_t)l->res & 15) == 0, "bad residual layout");
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) return nrrt_fail(NRRT_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available from the driver");
    const bool three_d = l->dims == 3;
    const int N = l->n, iD = three_d ? l->in_d : 1, iH = l->in_h, iW = l->in_w;
    NRRT_REQUIRE(N >= 1 && iD >= 1 && iH >= 1 && iW >= 1, "bad input dims");

    ConvParams& p = out.p;
    p = ConvParams{};
    int stride = 1, ksz = 3, pad = 1, up = 1;
    switch (l->kind) {
        case NRRT_CONV_K3S1: break;
        case NRRT_CONV_K3S2: stride = 2; break;
        case NRRT_CONV_K1S2: stride = 2; ksz = 1; pad = 0; break;
        case NRRT_CONV_K1S1: ksz = 1; pad = 0; break;
        case NRRT_CONVT_K2S2: ksz = 1; pad = 0; up = 2; break;
        case NRRT_CONV_UPFOLD: break;  // tile space = the low-resolution input grid, geometry of a 3x3 stride-1 conv
    }
    const bool fold = l->kind == NRRT_CONV_UPFOLD;
    if (fold)
        NRRT_REQUIRE(!three_d && l->cout % 128 == 0 && l->cin2 == 0 && l->res && l->out && !l->logits && !l->post_scale &&
                         l->res_coffset % 64 == 0 && !l->force_per_tap && !l->no_tma_store,
                     "folded ConvTranspose+conv layers: 2D, Cout multiple of 128, with a residual and a stored output");
    auto out_dim = [&](int i) { return (i + 2 * pad - ksz) / stride + 1; };
    p.N = N;
    p.W = out_dim(iW); p.H = out_dim(iH); p.D = three_d ? out_dim(iD) : 1;
    p.stride = stride; p.stride_d = three_d ? stride : 1;
    pick_box(p.W, p.H, p.D, p.bw, p.bh, p.bd);
    p.rows = p.bw * p.bh * p.bd;
    p.tiles_w = (p.W + p.bw - 1) / p.bw; p.tiles_h = (p.H + p.bh - 1) / p.bh; p.tiles_d = (p.D + p.bd - 1) / p.bd;
    const long long m_tiles = (long long)p.tiles_w * p.tiles_h * p.tiles_d * N;
    NRRT_REQUIRE(m_tiles < (1ll << 30), "too many tiles");
    p.m_tiles = (int)m_tiles;
    p.ntaps = 0;
    for (int dz = 0; dz < (three_d ? ksz : 1); ++dz)
        for (int dy = 0; dy < ksz; ++dy)
            for (int dx = 0; dx < ksz; ++dx) {
                p.tap[p.ntaps][0] = (int8_t)(dx - pad); p.tap[p.ntaps][1] = (int8_t)(dy - pad);
                p.tap[p.ntaps][2] = (int8_t)(three_d ? dz - pad : 0); p.tap[p.ntaps][3] = 0;
                ++p.ntaps;
            }
    const int cin_total = l->cin + l->cin2;
    p.cin_chunks = cin_total / 64;
    p.split_chunk = l->cin / 64;
    p.idx1 = l->in_index; p.idx2 = l->in2_index;
    const int groups = up == 2 ? (three_d ? 8 : 4) : 1;
    const int rows_total = groups * l->cout;
    const int block_n = fold ? 128 : (rows_total % 256 == 0 ? 256 : (rows_total % 128 == 0 ? 128 : 64));
    NRRT_REQUIRE(rows_total % block_n == 0, "Cout (x groups) = %d must be a multiple of 64", rows_total);
    p.n_tiles = rows_total / block_n;
    p.out = (__half*)l->out; p.out_cstride = l->out_cstride; p.out_coffset = l->out_coffset;
    p.up = up; p.up_d = (up == 2 && three_d) ? 2 : 1;
    p.oW = p.W * up; p.oH = p.H * up; p.oD = p.D * p.up_d;
    p.fold = fold ? 1 : 0;
    p.group = l->res_group > 1 ? l->res_group : 1;
    const int ups = fold ? 2 : 1;  // output grid of a folded layer = 2 x the tile-space grid (p.up stays 1: no weight groups)
    if (fold) { p.oW = 2 * p.W; p.oH = 2 * p.H; }
    p.cout = l->cout;
    p.scale = l->scale; p.shift = l->shift;
    p.scale2 = l->post_scale; p.shift2 = l->post_shift;
    NRRT_REQUIRE((l->post_scale == nullptr) == (l->post_shift == nullptr), "post_scale and post_shift go together");
    p.res = (const __half*)l->res; p.res_cstride = l->res_cstride; p.res_coffset = l->res_coffset;
    p.res_idx = l->res ? l->res_index : nullptr;
    p.relu = l->relu;
    p.poly = l->poly;
    p.poly_cases = l->poly_cases;
    p.inv_h = p.H > 1 ? 1.0f / (float)(p.H - 1) : 0.f;
    p.inv_w = p.W > 1 ? 1.0f / (float)(p.W - 1) : 0.f;
    if (l->poly) {
        NRRT_REQUIRE(!three_d, "the polynomial coordinate term is 2D only");
        NRRT_REQUIRE(((uintptr_t)l->poly & 15) == 0 && l->cout % 4 == 0, "poly must be 16-byte aligned");
        NRRT_REQUIRE(l->poly_cases == (up == 2 ? groups : 9), "poly_cases must be 9 (3x3 conv) or the number of weight groups (ConvTranspose)");
    }
    p.a_bytes = (uint32_t)p.rows * 128u;
    p.b_bytes = (uint32_t)block_n * 128u;
    p.head_w = l->head_weight; p.head_b = l->head_bias; p.logits = l->logits;
    p.poly3 = l->poly3;
    p.inv_d3 = p.D > 1 ? 1.0f / (float)(p.D - 1) : 0.f;
    p.inv_h3 = p.H > 1 ? 1.0f / (float)(p.H - 1) : 0.f;
    if (l->logits) {
        NRRT_REQUIRE(l->head_weight && p.n_tiles == 1 && up == 1, "fused head needs head_weight and all output channels in one N tile");
    } else {
        NRRT_REQUIRE(l->out, "null output pointer");
    }
    // 2D 3x3 stride-1 layers with a narrow N are L2-bound in the per-tap kernel: use the slab (halo-reuse) kernel
    // (3D: per depth slice, when the 16x16 tiles fill the slice and the pair kernel runs it)
    const bool halo = fold || (l->kind == NRRT_CONV_K3S1 && p.n_tiles == 1 && block_n <= 128 && !l->force_per_tap &&
                      (!three_d || (l->force_per_tap != 2 && p.W % 16 == 0 && p.H % 16 == 0)));
    const bool pair = l->force_per_tap != 2 && (halo || block_n >= 128);  // CTA pairs (cta_group::2)
    p.outer = fold ? 2 : (three_d ? 9 : 3);
    p.inner = fold ? 2 : 3;
    if (halo) {
        p.bw = 16; p.bh = 16; p.bd = 1; p.rows = 256;
        p.tiles_w = (p.W + 15) / 16; p.tiles_h = (p.H + 15) / 16; p.tiles_d = p.D;
        const long long mt = (long long)p.tiles_w * p.tiles_h * p.tiles_d * N;
        NRRT_REQUIRE(mt < (1ll << 30), "too many tiles");
        p.m_tiles = (int)mt;
    }

    // activation maps: dims (C, W, H, D, N) over the caller's NDHWC buffers (channel slices of cstride-wide rows)
    CUtensorMap &ma = out.ma, &ma2 = out.ma2, &mb = out.mb, &mr = out.mr, &mo = out.mo;
    auto make_act_map = [&](CUtensorMap* m, const void* ptr, int cin, int64_t cstride, int n_samples) -> CUresult {
        const cuuint64_t gdim[5] = {(cuuint64_t)cin, (cuuint64_t)iW, (cuuint64_t)iH, (cuuint64_t)iD, (cuuint64_t)n_samples};
        const cuuint64_t cs = (cuuint64_t)cstride * 2;
        const cuuint64_t gstr[4] = {cs, cs * iW, cs * iW * iH, cs * iW * iH * iD};
        cuuint32_t box[5] = {64, (cuuint32_t)(p.bw * stride), (cuuint32_t)(p.bh * stride), (cuuint32_t)(p.bd * p.stride_d), 1};
        if (halo) { box[1] = 8; box[2] = 18; box[3] = 1; }  // one slab: 18 pixel rows x 8 pixels
        const cuuint32_t estr[5] = {1, (cuuint32_t)stride, (cuuint32_t)stride, (cuuint32_t)p.stride_d, 1};
        return encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(ptr), gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    {
        CUresult r = make_act_map(&ma, l->in, l->cin, l->in_cstride, l->in_n > 0 ? l->in_n : N);
        if (r == CUDA_SUCCESS && l->cin2 > 0) r = make_act_map(&ma2, l->in2, l->cin2, l->in2_cstride, l->in2_n);
        else ma2 = ma;
        if (r != CUDA_SUCCESS) return nrrt_fail(NRRT_ERR_CUDA, "cuTensorMapEncodeTiled(activations) failed: %d (box %d,%d,%d stride %d)", (int)r, p.bw, p.bh, p.bd, stride);
    }
    {
        // folded: [4 parities][2x2 taps][Cin], then the 128 identity columns that carry the residual (p.res_mma)
        const cuuint64_t ktot = (cuuint64_t)(fold ? 16 : p.ntaps) * cin_total + (fold ? 128 : 0);
        const cuuint64_t gdim[2] = {ktot, (cuuint64_t)rows_total};
        const cuuint64_t gstr[1] = {ktot * 2};
        const cuuint32_t box[2] = {64, (cuuint32_t)(pair ? block_n / 2 : block_n)};  // a CTA of a pair loads half of B
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = encode(&mb, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(l->weight), gdim, gstr, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return nrrt_fail(NRRT_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
    }
    // pair kernels stage the residual tile through shared memory with TMA (output grid, 64-channel pieces)
    const bool tma_res = pair && l->res && up == 1 && !l->logits && l->cout % 64 == 0 && l->res_coffset % 64 == 0;
    // TMA-store epilogue: pair kernels with a stored output of whole 64-channel pieces (N = 128 slab kernel for now)
    p.ts = (pair && block_n >= 128 && !l->logits && l->out && l->cout % 64 == 0 && !l->no_tma_store) ? 1 : 0;
    mo = ma;
    if (p.ts) {
        const cuuint64_t gdim[5] = {(cuuint64_t)l->out_cstride, (cuuint64_t)p.oW, (cuuint64_t)p.oH, (cuuint64_t)p.oD, (cuuint64_t)N};
        const cuuint64_t cs = (cuuint64_t)l->out_cstride * 2;
        const cuuint64_t gstr[4] = {cs, cs * p.oW, cs * p.oW * p.oH, cs * p.oW * p.oH * p.oD};
        cuuint32_t box[5] = {64, (cuuint32_t)(p.bw * up), (cuuint32_t)(p.bh * up), (cuuint32_t)(p.bd * p.up_d), 1};
        if (halo) { box[1] = 8 * ups; box[2] = 16 * ups; box[3] = 1; }
        const cuuint32_t estr[5] = {1, (cuuint32_t)(up * ups), (cuuint32_t)(up * ups), (cuuint32_t)p.up_d, 1};
        const CUresult r = encode(&mo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, l->out, gdim, gstr, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return nrrt_fail(NRRT_ERR_CUDA, "cuTensorMapEncodeTiled(output) failed: %d", (int)r);
    }
    // interior polynomial through the tensor core: 2D 3x3 convs on the pair kernels with a staged residual (conv7.0 / conv8.0)
    p.pm = (tma_res && l->poly && l->poly_cases == 9 && (p.n_tiles == 1 || fold) && !l->no_poly_mma && !three_d && !l->post_scale &&
            ((halo && block_n == 128) || (!halo && block_n == 256 && l->kind == NRRT_CONV_K3S1 && p.rows == 128)) &&
            p.H > p.bh && p.W > p.bw) ? 1 : 0;  // (a tile holds at most one frame row and one frame column)
    mr = ma;
    if (tma_res) {
        const int rn = l->res_n > 0 ? l->res_n : N;
        const cuuint64_t gdim[5] = {(cuuint64_t)l->res_cstride, (cuuint64_t)p.oW, (cuuint64_t)p.oH, (cuuint64_t)p.oD, (cuuint64_t)rn};
        const cuuint64_t cs = (cuuint64_t)l->res_cstride * 2;
        const cuuint64_t gstr[4] = {cs, cs * p.oW, cs * p.oW * p.oH, cs * p.oW * p.oH * p.oD};
        cuuint32_t box[5] = {64, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bd, 1};
        if (halo) { box[1] = 8 * ups; box[2] = 16 * ups; box[3] = 1; }  // one column block of the 16x16 tile
        const cuuint32_t estr[5] = {1, (cuuint32_t)ups, (cuuint32_t)ups, 1, 1};
        const CUresult r = encode(&mr, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void*>(l->res), gdim, gstr, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return nrrt_fail(NRRT_ERR_CUDA, "cuTensorMapEncodeTiled(residual) failed: %d", (int)r);
    }
    // folded layers: the residual enters through the tensor core (conv_halo2_kernel, p.res_mma), not the epilogue
    // experiments (scripts/conv_probe.py): NRRT_CONV_DBG bits 1 = no TMA stores, 2 = no epilogue arithmetic, 8 = no polynomial,
    // 16 = no residual (folded layers); results are wrong with any of them set
    // Compiled out of the product build: a stray environment variable must not be able to corrupt every forward.
#ifdef NRRT_DEBUG_SWITCHES
    static const int dbg = getenv("NRRT_CONV_DBG") ? atoi(getenv("NRRT_CONV_DBG")) : 0;
#else
    constexpr int dbg = 0;
#endif
    p.dbg = dbg;
    if (fold && (dbg & 8)) { p.pm = 0; p.poly = nullptr; }
    const bool no_res = fold && (dbg & 16);
    if (no_res) p.res = nullptr;
    // residual through the tensor core: +25 % MMA work on conv8.0 buys an epilogue that never waits for the residual; in
    // the power-capped steady state of bench.py the two variants tie (3435 vs 3447 plans/s), so the staged one (less
    // tensor work) is the default and this one is opt-in
    const bool res_mma = fold && l->res_mma && !no_res;
    const bool res_st = tma_res && !res_mma && !no_res;
    if (res_mma) { p.res_mma = block_n / 64; p.res_k0 = 16 * cin_total; p.res = nullptr; }
    {
        const double in_pix = (double)iW * iH * iD, out_pix = (double)p.oW * p.oH * p.oD;
        const double taps_eff = fold ? 4.0 : (double)p.ntaps;  // per output pixel; ConvTranspose: one tap per output pixel
        out.flops = 2.0 * N * out_pix * l->cout * taps_eff * cin_total;
        const double w_elems = (double)rows_total * ((fold ? 16.0 : (double)p.ntaps) * cin_total + (fold ? 128.0 : 0.0));
        double b = 2.0 * std::min(N, l->in_n > 0 ? l->in_n : N) * in_pix * l->cin + 2.0 * w_elems;
        if (l->cin2 > 0) b += 2.0 * std::min(N, l->in2_n > 0 ? l->in2_n : N) * in_pix * l->cin2;
        if (l->res) b += 2.0 * std::min(N, l->res_n > 0 ? l->res_n : N) * out_pix * l->cout;
        b += l->logits ? 4.0 * N * out_pix : 2.0 * N * out_pix * l->cout;
        out.bytes = b;
    }
    if (fold) NRRT_REQUIRE(pair && tma_res && p.ts, "folded layer: residual / output layout not eligible for the TMA-staged epilogue");
    if (l->poly3)
        NRRT_REQUIRE(three_d && halo && pair && block_n == 128 && p.n_tiles == 1 && res_st && p.ts && !l->poly && p.H >= 4 && p.H % 2 == 0 &&
                         ((uintptr_t)l->poly3 & 15) == 0,
                     "poly3: 3D 3x3x3 layer with 128 output channels, a residual and whole 16x16 tiles per depth slice");
    // weight tiles one pair iteration consumes: a layer whose count equals an instance's B ring keeps them resident (p.bres)
    const int k_tiles = (fold || l->res_mma) ? -1 : p.cin_chunks * p.ntaps;
    if (halo && pair) {
        if (block_n == 128 && p.ts) {
            if (p.poly3) { out.launch = launch_halo2<128, 3, 4, true, false, true, true>; return NRRT_OK; }
            if (res_st && p.pm) { out.launch = launch_halo2<128, 3, 4, true, true, true>; return NRRT_OK; }
            if (p.pm) { out.launch = launch_halo2<128, 3, 4, false, true, true>; return NRRT_OK; }
            if (res_st) { out.launch = launch_halo2<128, 3, 5, true, false, true>; return NRRT_OK; }
            { out.launch = launch_halo2<128, 3, 5, false, false, true>; return NRRT_OK; }
        }
        if (res_st) {
            if (block_n == 128) { out.launch = launch_halo2<128, 3, 5, true>; return NRRT_OK; }
            {
                if (k_tiles == 9) { p.bres = 1; out.launch = launch_halo2<64, 4, 9, true>; return NRRT_OK; }
                out.launch = launch_halo2<64, 4, 8, true>; return NRRT_OK;
            }
        }
        if (block_n == 128) { out.launch = launch_halo2<128, 4, 6, false>; return NRRT_OK; }
        {
            if (k_tiles == 9) { p.bres = 1; out.launch = launch_halo2<64, 4, 9, false>; return NRRT_OK; }
            if (k_tiles == 18) { p.bres = 1; out.launch = launch_halo2<64, 4, 18, false>; return NRRT_OK; }
            out.launch = launch_halo2<64, 4, 8, false>; return NRRT_OK;
        }
    }
    if (halo) {
        if (block_n == 128) { out.launch = launch_halo<128, 4, 4>; return NRRT_OK; }
        { out.launch = launch_halo<64, 4, 8>; return NRRT_OK; }
    }
    if (pair) {
        if (p.ts) {
            if (block_n == 256) {
                if (res_st && p.pm) { out.launch = launch_conv2<256, 4, true, true, true>; return NRRT_OK; }
                if (res_st) { out.launch = launch_conv2<256, 4, true, false, true>; return NRRT_OK; }
                { out.launch = launch_conv2<256, 4, false, false, true>; return NRRT_OK; }
            }
            if (res_st) { out.launch = launch_conv2<128, 6, true, false, true>; return NRRT_OK; }
            { out.launch = launch_conv2<128, 6, false, false, true>; return NRRT_OK; }
        }
        if (res_st) {
            if (block_n == 256) { out.launch = launch_conv2<256, 4, true>; return NRRT_OK; }
            { out.launch = launch_conv2<128, 6, true>; return NRRT_OK; }
        }
        if (block_n == 256) { out.launch = launch_conv2<256, 6, false>; return NRRT_OK; }
        { out.launch = launch_conv2<128, 8, false>; return NRRT_OK; }
    }
    if (block_n == 256) { out.launch = launch_conv<256, 4>; return NRRT_OK; }
    if (block_n == 128) { out.launch = launch_conv<128, 6>; return NRRT_OK; }
    { out.launch = launch_conv<64, 8>; return NRRT_OK; }
}

}  // namespace
