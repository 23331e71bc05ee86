// conv_aux.cu — the small CUDA-core pieces of the CNN forward that are not GEMM-shaped: the 3-/1-channel stem
// convolution, the coordinate branch (4 tiny convs on a 2x2 / 2x3 grid), its align_corners interpolation into the
// concat buffers, and the 1x1 output head.  (train.py:118-149,171-203; ThreeD_train.py:112-147,169-204.)
#include <cuda_fp16.h>

#include "nrrt_common.cuh"

namespace {

// ---- stem: direct 3x3(x3) convolution, Cin <= 4, fp32 NCHW in -> fp16 channels-last out, bias + ReLU --------------
template <int DIMS>
__global__ void __launch_bounds__(256) stem_kernel(const float* __restrict__ x, const uint8_t* __restrict__ occ,
                                                   const int32_t* __restrict__ t2m, int n, int cin, int D, int H, int W,
                                                   const float* __restrict__ wgt, const float* __restrict__ bias, int cout,
                                                   int cout_pad, __half* __restrict__ out) {
    extern __shared__ float s_w[];  // [cout][cin*taps] + bias[cout]
    constexpr int TAPS = DIMS == 3 ? 27 : 9;
    const int kk = cin * TAPS;
    for (int i = threadIdx.x; i < cout * kk; i += blockDim.x) s_w[i] = wgt[i];
    for (int i = threadIdx.x; i < cout; i += blockDim.x) s_w[cout * kk + i] = bias[i];
    __syncthreads();
    const long long pixels = (long long)n * D * H * W;
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < pixels; pix += (long long)gridDim.x * blockDim.x) {
        long long t = pix;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H); t /= H;
        const int d = (int)(t % D);
        const int b = (int)(t / D);
        float in[4 * TAPS];
        const uint8_t* omap = occ ? occ + (size_t)__ldg(t2m + b) * D * H * W : nullptr;
        for (int c = 0; c < cin; ++c) {
            int q = 0;
            for (int dz = (DIMS == 3 ? -1 : 0); dz <= (DIMS == 3 ? 1 : 0); ++dz)
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx, ++q) {
                        const int zz = d + dz, yy = h + dy, xx = w + dx;
                        const bool ok = zz >= 0 && zz < D && yy >= 0 && yy < H && xx >= 0 && xx < W;
                        float v = 0.f;
                        if (ok) {
                            // map mode: every input channel is the {0,1}-normalised occupancy map (train.py:41-44)
                            // 3D: the map file is indexed [x][y][z] and the dataset's ToTensorV2 moves its LAST axis first
                            // (ThreeD_train.py:59-66), so network voxel (d, h, w) is map cell [x = h][y = w][z = d]
                            if (omap) v = __ldg(omap + (DIMS == 3 ? ((long long)yy * W + xx) * D + zz : ((long long)zz * H + yy) * W + xx)) != 0 ? 1.f : 0.f;
                            else v = __ldg(x + ((((long long)b * cin + c) * D + zz) * H + yy) * W + xx);
                        }
                        in[c * TAPS + q] = v;
                    }
        }
        __half* o = out + pix * cout_pad;
        for (int co = 0; co < cout; co += 8) {
            uint4 ov;
            __half2* hv = reinterpret_cast<__half2*>(&ov);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float a0 = s_w[cout * kk + co + 2 * e], a1 = s_w[cout * kk + co + 2 * e + 1];
                for (int k = 0; k < kk; ++k) {
                    a0 = fmaf(in[k], s_w[(co + 2 * e) * kk + k], a0);
                    a1 = fmaf(in[k], s_w[(co + 2 * e + 1) * kk + k], a1);
                }
                hv[e] = __floats2half2_rn(fmaxf(a0, 0.f), fmaxf(a1, 0.f));
            }
            *reinterpret_cast<uint4*>(o + co) = ov;
        }
        for (int co = cout; co < cout_pad; co += 8) *reinterpret_cast<uint4*>(o + co) = make_uint4(0, 0, 0, 0);
    }
}

// ---- coordinate branch: 3x3 conv (padding 1) + ReLU on a (gh x gw) grid with gh, gw <= 3, four levels --------------
// kTB tasks per CTA share every weight read; thread = output channel (weights pre-transposed to [tap][cin][cout] so
// the warp's loads are coalesced); inputs of the kTB tasks sit in shared memory and are read as broadcasts.
constexpr int kTB = 4;
constexpr int kMaxP = 9;

__global__ void __launch_bounds__(256) coords_kernel(const float* __restrict__ coords, int B, int gh, int gw,
                                                     const float* w0, const float* b0, const float* w1, const float* b1,
                                                     const float* w2, const float* b2, const float* w3, const float* b3,
                                                     float* f0, float* f1, float* f2, float* f3) {
    __shared__ float s_in[kTB * kMaxP * 256];
    const int P = gh * gw;
    const int t0 = blockIdx.x * kTB;
    const int nt = min(kTB, B - t0);
    const float* wts[4] = {w0, w1, w2, w3};
    const float* bs[4] = {b0, b1, b2, b3};
    float* outs[4] = {f0, f1, f2, f3};
    const int cins[4] = {1, 64, 128, 256}, couts[4] = {64, 128, 256, 512};
    for (int l = 0; l < 4; ++l) {
        const int cin = cins[l], cout = couts[l];
        // stage the inputs of this CTA's tasks: [t][p][ci]
        const float* src = l == 0 ? coords + (size_t)t0 * P : outs[l - 1] + (size_t)t0 * P * cin;
        for (int i = threadIdx.x; i < nt * P * cin; i += blockDim.x) s_in[i] = src[i];
        __syncthreads();
        for (int co = threadIdx.x; co < cout; co += blockDim.x) {
            float acc[kTB][kMaxP];
#pragma unroll
            for (int t = 0; t < kTB; ++t)
#pragma unroll
                for (int pp = 0; pp < kMaxP; ++pp) acc[t][pp] = bs[l][co];
            for (int tap = 0; tap < 9; ++tap) {
                const int ky = tap / 3 - 1, kx = tap % 3 - 1;
                const float* wp = wts[l] + (size_t)tap * cin * cout + co;
                for (int ci = 0; ci < cin; ++ci) {
                    const float w = __ldg(wp + (size_t)ci * cout);
#pragma unroll
                    for (int pp = 0; pp < kMaxP; ++pp) {
                        if (pp >= P) break;
                        const int y = pp / gw + ky, x = pp % gw + kx;
                        if (y < 0 || y >= gh || x < 0 || x >= gw) continue;
                        const int sp = y * gw + x;
#pragma unroll
                        for (int t = 0; t < kTB; ++t) acc[t][pp] = fmaf(s_in[(t * P + sp) * cin + ci], w, acc[t][pp]);
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < kTB; ++t) {
                if (t >= nt) break;
#pragma unroll
                for (int pp = 0; pp < kMaxP; ++pp) {
                    if (pp >= P) break;
                    outs[l][((size_t)(t0 + t) * P + pp) * cout + co] = fmaxf(acc[t][pp], 0.f);
                }
            }
        }
        __syncthreads();  // this CTA's global writes are visible to its own threads; s_in is free again
    }
}

// One layer of the branch for SMALL batches (the training step: 3 tasks; a short last chunk of a planning batch): the fused
// kernel above has B / kTB CTAs, i.e. one CTA walking 1.6 M weights alone.  Here a thread owns ONE output (task, position,
// channel) and runs the same chain as coords_kernel -- bias, then fmaf over (tap, ci) in that order -- so the results are
// bit-identical to the fused kernel's: which kernel a batch size selects must never show in the plans (a sharded run
// equals the single-GPU run bit for bit).  The weight loads do not depend on the accumulator and pipeline under the chain.
constexpr int kCoB = 64, kTpB = 4;
__global__ void __launch_bounds__(kCoB * kTpB) coords_layer_kernel(const float* __restrict__ in, int B, int gh, int gw, int cin, int cout,
                                                                   const float* __restrict__ wt, const float* __restrict__ bias,
                                                                   float* __restrict__ out) {
    extern __shared__ float s_in[];     // [kTpB][P][cin]: the inputs of the tasks this block touches (at most kTpB distinct)
    const int P = gh * gw;
    const int co = blockIdx.x * kCoB + threadIdx.x % kCoB;
    const int slot = threadIdx.x / kCoB;
    const int tp0 = blockIdx.y * kTpB;                       // first (task, position) pair of the block
    const int tp = tp0 + slot;
    for (int i = threadIdx.x; i < kTpB * P * cin; i += blockDim.x) {
        const int sl = i / (P * cin), r = i - sl * P * cin;
        const int t = (tp0 + sl) / P;
        s_in[i] = t < B ? in[(size_t)t * P * cin + r] : 0.f;
    }
    __syncthreads();
    if (co >= cout || tp >= B * P) return;
    const int pp = tp % P;
    const float* x = s_in + slot * P * cin;
    float acc = bias[co];
    for (int tap = 0; tap < 9; ++tap) {
        const int y = pp / gw + tap / 3 - 1, xq = pp % gw + tap % 3 - 1;
        if (y < 0 || y >= gh || xq < 0 || xq >= gw) continue;
        const float* wp = wt + (size_t)tap * cin * cout + co;
        const float* xp = x + (y * gw + xq) * cin;
#pragma unroll 8
        for (int ci = 0; ci < cin; ++ci) acc = fmaf(xp[ci], __ldg(wp + (size_t)ci * cout), acc);
    }
    out[(size_t)tp * cout + co] = fmaxf(acc, 0.f);
}

// ---- align_corners interpolation of the (gh x gw [x 1]) feature into a channel slice (PyTorch's fp32 formulas) -------
// Output pixels are ordered (b, a0, a1, a2); a0 interpolates the gh axis, a1 the gw axis, a2 is constant
// (2D: (H, W, 1); 3D: `coords.unsqueeze(-1)` makes the source (2, 3, 1), so the last output axis has a single source cell).
__global__ void __launch_bounds__(256) coords_fill_kernel(const float* __restrict__ feat, int B, int gh, int gw, int C,
                                                          int A0, int A1, int A2, __half* __restrict__ out,
                                                          long long cstride, long long coffset) {
    const int cv = C / 8;  // 8 channels (16 bytes) per thread
    const long long total = (long long)B * A0 * A1 * A2 * cv;
    const float rh = A0 > 1 ? (float)(gh - 1) / (float)(A0 - 1) : 0.f;  // area_pixel_compute_scale, align_corners=True
    const float rw = A1 > 1 ? (float)(gw - 1) / (float)(A1 - 1) : 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % cv);
        const long long pix = i / cv;
        long long t = pix / A2;
        const int x = (int)(t % A1); t /= A1;
        const int y = (int)(t % A0);
        const int b = (int)(t / A0);
        const float h1r = rh * y, w1r = rw * x;
        const int h1 = (int)h1r, w1 = (int)w1r;
        const int h1p = h1 < gh - 1 ? 1 : 0, w1p = w1 < gw - 1 ? 1 : 0;
        const float h1l = h1r - h1, h0l = 1.f - h1l, w1l = w1r - w1, w0l = 1.f - w1l;
        const float* f00 = feat + ((size_t)b * gh * gw + (size_t)h1 * gw + w1) * C + c8 * 8;
        const float* f01 = f00 + (size_t)w1p * C;
        const float* f10 = f00 + (size_t)h1p * gw * C;
        const float* f11 = f10 + (size_t)w1p * C;
        float v00[8], v01[8], v10[8], v11[8];
#pragma unroll
        for (int q = 0; q < 2; ++q) {  // 16-byte loads: the four corner vectors of this thread's 8 channels
            *reinterpret_cast<float4*>(v00 + 4 * q) = __ldg(reinterpret_cast<const float4*>(f00) + q);
            *reinterpret_cast<float4*>(v01 + 4 * q) = __ldg(reinterpret_cast<const float4*>(f01) + q);
            *reinterpret_cast<float4*>(v10 + 4 * q) = __ldg(reinterpret_cast<const float4*>(f10) + q);
            *reinterpret_cast<float4*>(v11 + 4 * q) = __ldg(reinterpret_cast<const float4*>(f11) + q);
        }
        uint4 ov;
        __half2* hv = reinterpret_cast<__half2*>(&ov);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float r[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int c = e * 2 + q;
                r[q] = h0l * (w0l * v00[c] + w1l * v01[c]) + h1l * (w0l * v10[c] + w1l * v11[c]);
            }
            hv[e] = __floats2half2_rn(r[0], r[1]);
        }
        *reinterpret_cast<uint4*>(out + pix * cstride + coffset + c8 * 8) = ov;
    }
}

// ---- 1x1 head: C -> 1, fp32 logits --------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_kernel(const __half* __restrict__ in, long long pixels, int C, long long cstride,
                                                   const float* __restrict__ wgt, float bias, float* __restrict__ logits) {
    extern __shared__ float s_wh[];
    for (int i = threadIdx.x; i < C; i += blockDim.x) s_wh[i] = wgt[i];
    __syncthreads();
    for (long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x; pix < pixels; pix += (long long)gridDim.x * blockDim.x) {
        const uint4* ip = reinterpret_cast<const uint4*>(in + pix * cstride);
        float acc = bias;
        for (int c8 = 0; c8 < C / 8; ++c8) {
            const uint4 v = __ldg(ip + c8);
            const __half2* hv = reinterpret_cast<const __half2*>(&v);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __half22float2(hv[e]);
                acc = fmaf(f.x, s_wh[c8 * 8 + 2 * e], acc);
                acc = fmaf(f.y, s_wh[c8 * 8 + 2 * e + 1], acc);
            }
        }
        logits[pix] = acc;
    }
}

// ---- coordinate channels as a polynomial epilogue term (2D) ---------------------------------------------------------
// The interpolated coordinate feature is bilinear in the pixel position: c(y,x) = G0 + ty*G1 + tx*G2 + ty*tx*G3 with
// ty = y/(H-1), tx = x/(W-1) and G = corner differences of the 2x2 feature.  A 3x3 conv over it is again bilinear:
//   sum_tap W_tap c(y+dy, x+dx) = E0 + ty*E1 + tx*E2 + ty*tx*E3,
//   E0 = M00 G0 + a M10 G1 + b M01 G2 + ab M11 G3,  E1 = M00 G1 + b M01 G3,  E2 = M00 G2 + a M10 G3,  E3 = M00 G3,
// with the tap moments M_uv = sum_tap dy^u dx^v W_tap over the taps that fall inside the image (9 border cases) and
// a = 1/(H-1), b = 1/(W-1).  So the coordinate channels never enter the GEMM: K shrinks by 25-50 % in the four layers
// that consume them, and they are applied in fp32 instead of being rounded to fp16.
// moments: [cases][n_uv][Cc][Cout] (uv order 00, 10, 01, 11);  out E: [B][cases][4][Cout].
constexpr int kPB = 4;  // tasks per CTA

__global__ void __launch_bounds__(256) coords_poly_kernel(const float* __restrict__ feat, int B, int Cc,
                                                          const float* __restrict__ moments, int cases, int n_uv, int Cout,
                                                          float a, float b, float* __restrict__ E) {
    extern __shared__ float s_g[];  // [kPB][4][Cc]
    const int t0 = blockIdx.x * kPB;
    const int nt = min(kPB, B - t0);
    for (int i = threadIdx.x; i < nt * Cc; i += blockDim.x) {
        const int t = i / Cc, c = i % Cc;
        const float* f = feat + (size_t)(t0 + t) * 4 * Cc + c;  // corners p = y*2 + x
        const float f00 = f[0], f01 = f[Cc], f10 = f[2 * Cc], f11 = f[3 * Cc];
        float* g = s_g + (size_t)t * 4 * Cc + c;
        g[0] = f00; g[Cc] = f10 - f00; g[2 * Cc] = f01 - f00; g[3 * Cc] = f00 - f01 - f10 + f11;
    }
    __syncthreads();
    const float uvw[4][4] = {{1.f, 1.f, 1.f, 1.f}, {a, 0.f, a, 0.f}, {b, b, 0.f, 0.f}, {a * b, 0.f, 0.f, 0.f}};
    for (int co = threadIdx.x; co < Cout; co += blockDim.x) {
        for (int k = 0; k < cases; ++k) {
            float acc[kPB][4];
#pragma unroll
            for (int t = 0; t < kPB; ++t)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[t][e] = 0.f;
            for (int uv = 0; uv < n_uv; ++uv) {
                const float* mp = moments + ((size_t)(k * n_uv + uv) * Cc) * Cout + co;
                for (int ci = 0; ci < Cc; ++ci) {
                    const float w = __ldg(mp + (size_t)ci * Cout);
#pragma unroll
                    for (int t = 0; t < kPB; ++t) {
                        const float* g = s_g + (size_t)t * 4 * Cc + ci;
                        if (uv == 0) {
                            acc[t][0] = fmaf(w, g[0], acc[t][0]); acc[t][1] = fmaf(w, g[Cc], acc[t][1]);
                            acc[t][2] = fmaf(w, g[2 * Cc], acc[t][2]); acc[t][3] = fmaf(w, g[3 * Cc], acc[t][3]);
                        } else if (uv == 1) {  // M10: a*G1 -> E0, a*G3 -> E2
                            acc[t][0] = fmaf(w * uvw[1][0], g[Cc], acc[t][0]); acc[t][2] = fmaf(w * uvw[1][2], g[3 * Cc], acc[t][2]);
                        } else if (uv == 2) {  // M01: b*G2 -> E0, b*G3 -> E1
                            acc[t][0] = fmaf(w * uvw[2][0], g[2 * Cc], acc[t][0]); acc[t][1] = fmaf(w * uvw[2][1], g[3 * Cc], acc[t][1]);
                        } else {               // M11: ab*G3 -> E0
                            acc[t][0] = fmaf(w * uvw[3][0], g[3 * Cc], acc[t][0]);
                        }
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < kPB; ++t) {
                if (t >= nt) break;
#pragma unroll
                for (int e = 0; e < 4; ++e) E[(((size_t)(t0 + t) * cases + k) * 4 + e) * Cout + co] = acc[t][e];
            }
        }
    }
}

// ---- encoder reuse: copy a channel slice of per-map features into every task's concat buffer ----------------------
__global__ void __launch_bounds__(256) gather_channels_kernel(const __half* __restrict__ src, long long src_cstride,
                                                              long long src_coffset, __half* __restrict__ dst,
                                                              long long dst_cstride, long long dst_coffset, int C,
                                                              long long pixels, const int32_t* __restrict__ index, int B) {
    const int cv = C / 8;
    const long long total = (long long)B * pixels * cv;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % cv);
        const long long t = i / cv;
        const long long pix = t % pixels;
        const int b = (int)(t / pixels);
        const long long sp = (long long)__ldg(index + b) * pixels + pix;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + sp * src_cstride + src_coffset + c8 * 8));
        *reinterpret_cast<uint4*>(dst + ((long long)b * pixels + pix) * dst_cstride + dst_coffset + c8 * 8) = v;
    }
}


// ---- conv6.0 hoisted out of the per-task path (2D) ------------------------------------------------------------------
// conv6.0 is linear in torch.cat((upconv6(cat(x5, coords_x5)), x4, coords_x4)) (train.py:192-195), and only the
// coordinate channels depend on the task.  Its pre-activation therefore splits into a per-MAP tensor (tensor cores, once
// per map) plus a per-TASK field that is piecewise bilinear in the pixel position:
//   (a) the interpolated coords_x4 channels: the 9-border-case polynomial of nrrt_coords_poly (basis ty = Y/(H-1), tx = X/(W-1));
//   (b) conv6.0 over the coordinate part of upconv6's output.  That part is, per output parity g = (Y&1, X&1), bilinear
//       in the ConvTranspose input position (s, t) = (i/(H3-1), j/(W3-1)), (i, j) = (Y>>1, X>>1): U_g = sum_k Eu[g][k] phi_k(s,t),
//       phi = (1, s, t, st).  A 3x3 tap (dy, dx) of output pixel (Y, X) with parity (a, b) reads parity
//       g' = ((a+dy)&1, (b+dx)&1) at (i + floor((a+dy)/2), j + floor((b+dx)/2)), so the tap sum is again bilinear, with
//       coefficients that depend on (a, b) and on which taps fall inside the image: 4 row classes (Y = 0, interior even,
//       interior odd, Y = H-1) x 4 column classes = 16 cases.
// sgemm_kernel computes V[b][g][k][tap][co] = sum_c Eu[b][g][k][c] * Wtap[c][tap][co] in fp32 (the coordinate features are
// O(100) and are never rounded to fp16); poly16_combine_kernel shifts/sums the taps per case, converts to the (ty, tx)
// basis, adds (a) and the conv bias: P16[b][case][4][co].  poly_relu_kernel then forms the layer's activation
// relu(pre[map] + P16(task)) for every task.
constexpr int GM = 128, GN = 128, GK = 16;

__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                    float* __restrict__ Cm, int M, int N, int K) {
    __shared__ __align__(16) float sA[GK][GM + 4];  // A tile transposed: [k][m]
    __shared__ __align__(16) float sB[GK][GN];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    // global -> register staging: A 128 rows x 16 k = 512 float4 (2 per thread), B 16 k x 128 n = 512 float4 (2 per thread)
    float4 ra[2], rb[2];
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int idx = tid + q * 256;
            const int row = idx >> 2, k4 = idx & 3;
            ra[q] = (m0 + row < M) ? __ldg(reinterpret_cast<const float4*>(A + (size_t)(m0 + row) * K + k0 + k4 * 4))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
            const int kr = idx >> 5, n4 = idx & 31;
            rb[q] = __ldg(reinterpret_cast<const float4*>(Bm + (size_t)(k0 + kr) * N + n0 + n4 * 4));
        }
    };
    load_tiles(0);
    for (int k0 = 0; k0 < K; k0 += GK) {
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int idx = tid + q * 256;
            const int row = idx >> 2, k4 = idx & 3;
            sA[k4 * 4 + 0][row] = ra[q].x; sA[k4 * 4 + 1][row] = ra[q].y;
            sA[k4 * 4 + 2][row] = ra[q].z; sA[k4 * 4 + 3][row] = ra[q].w;
            const int kr = idx >> 5, n4 = idx & 31;
            *reinterpret_cast<float4*>(&sB[kr][n4 * 4]) = rb[q];
        }
        __syncthreads();
        if (k0 + GK < K) load_tiles(k0 + GK);
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&sA[kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&sA[kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&sB[kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&sB[kk][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
        if (row >= M) continue;
        float* cp = Cm + (size_t)row * N + n0;
        *reinterpret_cast<float4*>(cp + tx * 4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        *reinterpret_cast<float4*>(cp + 64 + tx * 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
}

// V: [B][4 g][4 k][9 tap][Cout];  e_conv: [B][9][4][Cout] (basis ty, tx of the conv's grid);  P16: [B][16][4][Cout].
// a3 = 1/(H3-1), b3 = 1/(W3-1) of the ConvTranspose input grid; ly = (H-1)/(2(H3-1)), lx likewise: s = ly*ty - a*a3/2.
__global__ void __launch_bounds__(128) poly16_combine_kernel(const float* __restrict__ V, const float* __restrict__ e_conv,
                                                             const float* __restrict__ bias, int B, int Cout, float a3, float b3,
                                                             float ly, float lx, float* __restrict__ P16) {
    const int co = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (co >= Cout || b >= B) return;
    float acc[16][4];
#pragma unroll
    for (int cs = 0; cs < 16; ++cs)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[cs][e] = 0.f;
    const float* vb = V + (size_t)b * 16 * 9 * Cout + co;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int tap = (dy + 1) * 3 + (dx + 1);
            float v[4][4];  // [g][k]
#pragma unroll
            for (int g = 0; g < 4; ++g)
#pragma unroll
                for (int k = 0; k < 4; ++k) v[g][k] = __ldg(vb + ((size_t)(g * 4 + k) * 9 + tap) * Cout);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if ((r == 0 && dy < 0) || (r == 3 && dy > 0)) continue;
                const int a = r >= 2 ? 1 : 0;
                const int ap = (a + dy) & 1;
                const int sy = (a + dy) < 0 ? -1 : ((a + dy) >> 1);  // floor((a+dy)/2)
                const float my = a3 * ((float)sy - 0.5f * (float)a);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if ((c == 0 && dx < 0) || (c == 3 && dx > 0)) continue;
                    const int bb = c >= 2 ? 1 : 0;
                    const int bp = (bb + dx) & 1;
                    const int sx = (bb + dx) < 0 ? -1 : ((bb + dx) >> 1);
                    const float mx = b3 * ((float)sx - 0.5f * (float)bb);
                    const int g = ap * 2 + bp;
                    float* o = acc[r * 4 + c];
                    o[0] += v[g][0] + my * v[g][1] + mx * v[g][2] + (my * mx) * v[g][3];
                    o[1] += ly * (v[g][1] + mx * v[g][3]);
                    o[2] += lx * (v[g][2] + my * v[g][3]);
                    o[3] += (ly * lx) * v[g][3];
                }
            }
        }
    const float bi = __ldg(bias + co);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int c9 = (r == 0 ? 0 : (r == 3 ? 2 : 1)) * 3 + (c == 0 ? 0 : (c == 3 ? 2 : 1));
            const float* ec = e_conv + (((size_t)b * 9 + c9) * 4) * Cout + co;
            float* o = P16 + (((size_t)b * 16 + r * 4 + c) * 4) * Cout + co;
            o[0] = acc[r * 4 + c][0] + __ldg(ec) + bi;
            o[(size_t)Cout] = acc[r * 4 + c][1] + __ldg(ec + Cout);
            o[(size_t)2 * Cout] = acc[r * 4 + c][2] + __ldg(ec + 2 * (size_t)Cout);
            o[(size_t)3 * Cout] = acc[r * 4 + c][3] + __ldg(ec + 3 * (size_t)Cout);
        }
}

// out[b][Y][X][:] = fp16(relu(pre[index[b]][Y][X][:] + P0 + ty*P1 + tx*P2 + ty*tx*P3)), P = P16[b][case(Y, X)].
// One CTA per (row Y, task b): 256 threads = (C/8 channel groups) x pixel lanes; a lane walks X = lane, lane + lanes, ...
// (lanes is even, so the column parity and hence the interior coefficient set of a thread is fixed and lives in registers).
__global__ void __launch_bounds__(256) poly_relu_kernel(const __half* __restrict__ pre, const int32_t* __restrict__ index,
                                                        const float* __restrict__ P16, __half* __restrict__ out, int H, int W,
                                                        int C, float inv_h, float inv_w) {
    const int cgs = C >> 3, lanes = 256 / cgs;
    const int cg = threadIdx.x % cgs, lane = threadIdx.x / cgs;
    const int Y = blockIdx.x, b = blockIdx.y;
    const int r = Y == 0 ? 0 : (Y == H - 1 ? 3 : 1 + (Y & 1));
    const float ty = (float)Y * inv_h;
    const int src = index ? __ldg(index + b) : b;
    const __half* prow = pre + (((size_t)src * H + Y) * W) * C + cg * 8;
    __half* orow = out + (((size_t)b * H + Y) * W) * C + cg * 8;
    const float* pb = P16 + ((size_t)b * 16 + r * 4) * 4 * C + cg * 8;
    // row-reduced coefficients of a column class: value = A + tx * Bc
    auto load_class = [&](int c, float (&A)[8], float (&Bc)[8]) {
        const float* q = pb + (size_t)c * 4 * C;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float4 p0 = __ldg(reinterpret_cast<const float4*>(q + h * 4));
            const float4 p1 = __ldg(reinterpret_cast<const float4*>(q + C + h * 4));
            const float4 p2 = __ldg(reinterpret_cast<const float4*>(q + 2 * C + h * 4));
            const float4 p3 = __ldg(reinterpret_cast<const float4*>(q + 3 * C + h * 4));
            A[h * 4 + 0] = fmaf(ty, p1.x, p0.x); A[h * 4 + 1] = fmaf(ty, p1.y, p0.y);
            A[h * 4 + 2] = fmaf(ty, p1.z, p0.z); A[h * 4 + 3] = fmaf(ty, p1.w, p0.w);
            Bc[h * 4 + 0] = fmaf(ty, p3.x, p2.x); Bc[h * 4 + 1] = fmaf(ty, p3.y, p2.y);
            Bc[h * 4 + 2] = fmaf(ty, p3.z, p2.z); Bc[h * 4 + 3] = fmaf(ty, p3.w, p2.w);
        }
    };
    float Ai[8], Bi[8];
    load_class(1 + (lane & 1), Ai, Bi);
#pragma unroll 4
    for (int X = lane; X < W; X += lanes) {  // (unrolled: four 16-byte loads in flight per thread; the kernel is a copy stream)
        float Ae[8], Be[8];
        const bool edge = X == 0 || X == W - 1;
        if (edge) load_class(X == 0 ? 0 : 3, Ae, Be);
        const float tx = (float)X * inv_w;
        const uint4 pv = __ldg(reinterpret_cast<const uint4*>(prow + (size_t)X * C));
        const __half2* ph = reinterpret_cast<const __half2*>(&pv);
        uint4 ov;
        __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(ph[e]);
            const float a0 = edge ? Ae[2 * e] : Ai[2 * e], a1 = edge ? Ae[2 * e + 1] : Ai[2 * e + 1];
            const float b0 = edge ? Be[2 * e] : Bi[2 * e], b1 = edge ? Be[2 * e + 1] : Bi[2 * e + 1];
            const float v0 = fminf(fmaxf(f.x + fmaf(tx, b0, a0), 0.f), 65504.f);
            const float v1 = fminf(fmaxf(f.y + fmaf(tx, b1, a1), 0.f), 65504.f);
            oh[e] = __floats2half2_rn(v0, v1);
        }
        *reinterpret_cast<uint4*>(orow + (size_t)X * C) = ov;
    }
}

// ---- nrrt_coords_poly as a GEMM -------------------------------------------------------------------------------------
// The direct kernel above re-reads the whole moment tensor per 4 tasks (16 ms for conv6.0 at 4096 tasks).  The same
// contraction is C[(b, g)][(case, uv, co)] = sum_c G[b][g][c] * M[c][(case, uv, co)] with G = corner differences of the
// coordinate feature (4 vectors per task): one pass of sgemm_kernel, then E is a few scaled sums of C entries.
__global__ void __launch_bounds__(256) coords_g_kernel(const float* __restrict__ feat, long long total, int Cc, float* __restrict__ G) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / Cc;
        const int c = (int)(i - b * Cc);
        const float* f = feat + b * 4 * Cc + c;  // corners p = y*2 + x
        const float f00 = f[0], f01 = f[Cc], f10 = f[2 * Cc], f11 = f[3 * Cc];
        float* g = G + b * 4 * Cc + c;
        g[0] = f00; g[Cc] = f10 - f00; g[2 * Cc] = f01 - f00; g[3 * Cc] = f00 - f01 - f10 + f11;
    }
}
// Cw: [B][4 g][cases][n_uv][Cout];  E: [B][cases][4][Cout]
__global__ void __launch_bounds__(256) coords_combine_kernel(const float* __restrict__ Cw, long long total, int cases, int n_uv,
                                                             int Cout, float a, float b, float* __restrict__ E) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % Cout);
        const long long t = i / Cout;
        const int k = (int)(t % cases);
        const long long bb = t / cases;
        const size_t gs = (size_t)cases * n_uv * Cout;  // stride between the four G rows of a task
        const float* c0 = Cw + (size_t)bb * 4 * gs + ((size_t)k * n_uv) * Cout + co;
        float e0, e1, e2, e3;
        if (n_uv == 1) {
            e0 = c0[0]; e1 = c0[gs]; e2 = c0[2 * gs]; e3 = c0[3 * gs];
        } else {  // uv order 00, 10, 01, 11
            const size_t u = (size_t)Cout;
            e0 = c0[0] + a * c0[gs + u] + b * c0[2 * gs + 2 * u] + (a * b) * c0[3 * gs + 3 * u];
            e1 = c0[gs] + b * c0[3 * gs + 2 * u];
            e2 = c0[2 * gs] + a * c0[3 * gs + u];
            e3 = c0[3 * gs];
        }
        float* e = E + (((size_t)bb * cases + k) * 4) * Cout + co;
        e[0] = e0; e[Cout] = e1; e[2 * (size_t)Cout] = e2; e[3 * (size_t)Cout] = e3;
    }
}

// ---- 3D: coordinate channels as a post-pass polynomial ---------------------------------------------------------------
// ThreeD_UNet_cooler interpolates the (2, 3) coordinate feature trilinearly to (D, H, W) (ThreeD_train.py:180-191): linear
// in d, PIECEWISE linear in h (3 points: a knot in the middle, which falls between two voxels for even H), constant in w.
// Per h-piece p the field is bilinear in (td, th) = (d/(D-1), h/(H-1)):  A0 + td*A1 + th*A2 + td*th*A3  with
//   G0 = f[0][p], G1 = f[1][p]-f[0][p], G2 = f[0][p+1]-f[0][p], G3 = f[0][p]-f[0][p+1]-f[1][p]+f[1][p+1]
//   A0 = G0 - p*G2, A1 = G1 - p*G3, A2 = 2*G2, A3 = 2*G3                      (source coordinate along h is 2*th).
// A 3x3x3 conv over it is again bilinear in (td, th) per (d-border, h-class, w-border) case, 3 x 6 x 3 = 54 cases (h-class:
// first row, inside piece 0, last row of piece 0, first row of piece 1, inside piece 1, last row), with tap moments per
// piece.  Stacking the two pieces along the channel axis turns this into the 2D contraction: nrrt_coords_poly_gemm with
// Cc' = 2*Cc, cases = 54 and G given directly.  The conv itself runs without the coordinate channels (K -33...-50 %) and
// without ReLU; poly_relu3d_kernel adds the field (a per-row constant: it does not depend on w) and applies the ReLU.
__global__ void __launch_bounds__(256) coords_pieces_kernel(const float* __restrict__ feat, long long total, int Cc,
                                                            float* __restrict__ A) {
    // feat [B][6 = a*3 + b][Cc] -> A [B][4][2*Cc] (piece p in channels [p*Cc, (p+1)*Cc))
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / (2 * Cc);
        const int r = (int)(i - b * 2 * Cc), p = r / Cc, c = r - p * Cc;
        const float* f = feat + b * 6 * Cc + c;
        const float f0p = f[(size_t)p * Cc], f0q = f[(size_t)(p + 1) * Cc], f1p = f[(size_t)(3 + p) * Cc], f1q = f[(size_t)(4 + p) * Cc];
        const float g0 = f0p, g1 = f1p - f0p, g2 = f0q - f0p, g3 = f0p - f0q - f1p + f1q;
        float* a = A + b * 4 * 2 * Cc + r;
        a[0] = g0 - (float)p * g2;
        a[(size_t)2 * Cc] = g1 - (float)p * g3;
        a[(size_t)4 * Cc] = 2.f * g2;
        a[(size_t)6 * Cc] = 2.f * g3;
    }
}

// x[b][d][h][w][:] = fp16(relu(x + E0 + td*E1 + th*E2 + td*th*E3)), E = E[b][case(d, h, w)]: one CTA per (d, h) row and task.
__global__ void __launch_bounds__(256) poly_relu3d_kernel(__half* __restrict__ x, const float* __restrict__ E, int D, int H, int W,
                                                          int C, float inv_d, float inv_h) {
    const int cgs = C >> 3, lanes = 256 / cgs;
    const int cg = threadIdx.x % cgs, lane = threadIdx.x / cgs;
    const int d = blockIdx.x / H, h = blockIdx.x - d * H, b = blockIdx.y;
    const int dc = d == 0 ? 0 : (d == D - 1 ? 2 : 1);
    const int hh = H >> 1;
    const int hc = h == 0 ? 0 : (h == H - 1 ? 5 : (h < hh - 1 ? 1 : (h == hh - 1 ? 2 : (h == hh ? 3 : 4))));
    const float td = (float)d * inv_d, th = (float)h * inv_h;
    const float* eb = E + ((size_t)b * 54 + (size_t)(dc * 6 + hc) * 3) * 4 * C + cg * 8;
    __half* row = x + ((((size_t)b * D + d) * H + h) * W) * C + cg * 8;
    auto load_class = [&](int wc, float (&A)[8]) {
        const float* q = eb + (size_t)wc * 4 * C;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float4 e0 = __ldg(reinterpret_cast<const float4*>(q + k * 4));
            const float4 e1 = __ldg(reinterpret_cast<const float4*>(q + C + k * 4));
            const float4 e2 = __ldg(reinterpret_cast<const float4*>(q + 2 * C + k * 4));
            const float4 e3 = __ldg(reinterpret_cast<const float4*>(q + 3 * C + k * 4));
            A[k * 4 + 0] = e0.x + td * e1.x + th * (e2.x + td * e3.x);
            A[k * 4 + 1] = e0.y + td * e1.y + th * (e2.y + td * e3.y);
            A[k * 4 + 2] = e0.z + td * e1.z + th * (e2.z + td * e3.z);
            A[k * 4 + 3] = e0.w + td * e1.w + th * (e2.w + td * e3.w);
        }
    };
    float Ai[8];
    load_class(1, Ai);
    for (int w = lane; w < W; w += lanes) {
        float Ae[8];
        const bool edge = w == 0 || w == W - 1;
        if (edge) load_class(w == 0 ? 0 : 2, Ae);
        uint4 v = *reinterpret_cast<const uint4*>(row + (size_t)w * C);
        __half2* hv = reinterpret_cast<__half2*>(&v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(hv[e]);
            const float a0 = edge ? Ae[2 * e] : Ai[2 * e], a1 = edge ? Ae[2 * e + 1] : Ai[2 * e + 1];
            hv[e] = __floats2half2_rn(fminf(fmaxf(f.x + a0, 0.f), 65504.f), fminf(fmaxf(f.y + a1, 0.f), 65504.f));
        }
        *reinterpret_cast<uint4*>(row + (size_t)w * C) = v;
    }
}

int grid_for(long long work, int block) {
    long long g = (work + block - 1) / block;
    return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

// ---- nrrt_poly_fold: the epilogue polynomial of a 3x3 conv re-expressed for the folded ConvTranspose+conv layer -------
// (conv_tc.cu, conv_halo2_kernel, p.fold).  The conv's output pixel (h, w) = (2i + a, 2j + b) is pixel (i, j) of parity
// (a, b) in the folded layer's tile space, the (H, W) low-resolution grid.  With ty = h / (2H - 1) = al_h * ty' + be_h,
// ty' = i / (H - 1), al_h = 2 (H - 1) / (2H - 1), be_h = a / (2H - 1) (columns alike), the bilinear form
// E0 + ty E1 + tx E2 + ty tx E3 becomes E0' + ty' E1' + tx' E2' + ty' tx' E3' with the coefficients below.  Border cases:
// output row 0 exists only in parity a = 0 (tile row 0) and the last output row only in a = 1 (last tile row); the other
// tile-space frame rows of a parity are interior pixels of the conv and get the interior case (columns alike).
// `cases_const` [9][Cout] (optional) is added to E0: the ConvTranspose bias seen through the conv taps inside the image.
__global__ void poly_fold_kernel(const float* __restrict__ E, const float* __restrict__ cases_const, int B, int Cout, float al_h,
                                 float al_w, float inv_oh, float inv_ow, float* __restrict__ Ef) {
    const long long total = (long long)B * 36 * Cout;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cout);
        long long r = i / Cout;
        const int kc = (int)(r % 9); r /= 9;
        const int par = (int)(r % 4);
        const long long b = r / 4;
        const int a = par >> 1, bb = par & 1;
        const int rt = kc / 3, ct = kc % 3;
        const int ro = (rt == 0 && a == 0) ? 0 : ((rt == 2 && a == 1) ? 2 : 1);
        const int co = (ct == 0 && bb == 0) ? 0 : ((ct == 2 && bb == 1) ? 2 : 1);
        const float* src = E + ((b * 9 + ro * 3 + co) * 4) * Cout + c;
        float e0 = src[0];
        const float e1 = src[Cout], e2 = src[2 * (size_t)Cout], e3 = src[3 * (size_t)Cout];
        if (cases_const) e0 += __ldg(cases_const + (ro * 3 + co) * Cout + c);
        const float bh = (float)a * inv_oh, bw = (float)bb * inv_ow;
        float* dst = Ef + (((b * 4 + par) * 9 + kc) * 4) * Cout + c;
        dst[0] = e0 + bh * e1 + bw * e2 + bh * bw * e3;
        dst[Cout] = al_h * (e1 + bw * e3);
        dst[2 * (size_t)Cout] = al_w * (e2 + bh * e3);
        dst[3 * (size_t)Cout] = al_h * al_w * e3;
    }
}

// ---- stem_lut_kernel: the first conv fed from a binary 2D occupancy map as a table lookup -----------------------------
// Map mode in 2D: every input channel is the same {0,1} image (train.py:41-44), so the layer's output at a pixel depends only
// on the 9 bits of its 3x3 neighbourhood (zero padding = bit 0): relu(bias + sum over set taps of sum_c w[co][c][tap]).
// Each CTA builds the 512-entry table [pattern][cout] in shared memory (fp32 sums, fp16 entries: 32 KB for cout = 32) and
// then the kernel is a pure store stream: a warp takes 32 consecutive pixels, each lane computes its pixel's pattern, and
// the 32 x cout_pad fp16 outputs leave as fully coalesced 512-byte rows (lane l writes 16-byte chunk l % 8 of pixel
// 4 * step + l / 8, copied from the table or zero for the padding channels).  268 MB per 8 maps at 512^2: HBM-write bound
// (the per-pixel FMA version spent 0.51 ms per 8 maps, 10x the store time).
__global__ void __launch_bounds__(256) stem_lut_kernel(const uint8_t* __restrict__ occ, const int32_t* __restrict__ t2m, int n,
                                                       int cin, int H, int W, const float* __restrict__ wgt,
                                                       const float* __restrict__ bias, int cout, int cout_pad,
                                                       __half* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __half* lut = reinterpret_cast<__half*>(s_raw);                      // [512][cout]
    float* wsum = reinterpret_cast<float*>(s_raw + (size_t)512 * cout * 2);  // [cout][9] summed over input channels, + bias[cout]
    for (int i = threadIdx.x; i < cout * 9; i += blockDim.x) {
        const int co = i / 9, tap = i - co * 9;
        float a = 0.f;
        for (int c = 0; c < cin; ++c) a += wgt[(co * cin + c) * 9 + tap];
        wsum[i] = a;
    }
    for (int i = threadIdx.x; i < cout; i += blockDim.x) wsum[cout * 9 + i] = bias[i];
    __syncthreads();
    for (int i = threadIdx.x; i < 512 * cout; i += blockDim.x) {
        const int pat = i / cout, co = i - pat * cout;
        float a = wsum[cout * 9 + co];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
            if ((pat >> tap) & 1) a += wsum[co * 9 + tap];
        lut[i] = __float2half_rn(fmaxf(a, 0.f));
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long pixels = (long long)n * H * W;
    const long long n_groups = (pixels + 31) >> 5;
    const int chunks = cout_pad >> 3;  // 16-byte chunks per pixel
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long grp = warp0; grp < n_groups; grp += n_warps) {
        const long long pix = (grp << 5) + lane;
        int pat = 0;
        if (pix < pixels) {
            long long t = pix;
            const int w = (int)(t % W); t /= W;
            const int h = (int)(t % H);
            const int b = (int)(t / H);
            const uint8_t* omap = occ + (size_t)__ldg(t2m + b) * H * W;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const int yy = h + dy, xx = w + dx;
                    if (yy >= 0 && yy < H && xx >= 0 && xx < W && __ldg(omap + (long long)yy * W + xx) != 0)
                        pat |= 1 << ((dy + 1) * 3 + dx + 1);
                }
        }
        uint4* dst = reinterpret_cast<uint4*>(out + (grp << 5) * cout_pad);
        const int total = 32 * chunks;  // 16-byte chunks of this group of 32 pixels, contiguous in memory
        for (int i = lane; i < total; i += 32) {
            const int px = i / chunks, ch = i - px * chunks;
            const int ppat = __shfl_sync(0xffffffffu, pat, px);
            if ((grp << 5) + px < pixels) {
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (ch * 8 < cout) v = *reinterpret_cast<const uint4*>(lut + (size_t)ppat * cout + ch * 8);
                dst[i] = v;
            }
        }
    }
}

}  // namespace

static int stem_launch(int dims, const float* x, const uint8_t* occ, const int32_t* t2m, int n, int cin, int d, int h, int w,
                       const float* weight, const float* bias, int cout, int cout_pad, void* out, void* stream) {
    NRRT_REQUIRE((x || (occ && t2m)) && weight && bias && out, "null pointer");
    NRRT_REQUIRE((dims == 2 || dims == 3) && cin >= 1 && cin <= 4 && cout % 8 == 0 && cout_pad % 8 == 0 && cout_pad >= cout, "bad stem shape");
    if (dims == 2) d = 1;
    const int taps = dims == 3 ? 27 : 9;
    const size_t smem = ((size_t)cout * cin * taps + cout) * sizeof(float);
    const long long pixels = (long long)n * d * h * w;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t lut_smem = (size_t)512 * cout * 2 + ((size_t)cout * 9 + cout) * sizeof(float);
    if (dims == 2 && occ && lut_smem <= 48 * 1024) {  // binary map input: table lookup + coalesced store stream
        stem_lut_kernel<<<148 * 4, 256, lut_smem, st>>>(occ, t2m, n, cin, h, w, weight, bias, cout, cout_pad, (__half*)out);
        NRRT_CUDA_TRY(cudaGetLastError());
        return NRRT_OK;
    }
    if (dims == 2) stem_kernel<2><<<grid_for(pixels, 256), 256, smem, st>>>(x, occ, t2m, n, cin, d, h, w, weight, bias, cout, cout_pad, (__half*)out);
    else stem_kernel<3><<<grid_for(pixels, 256), 256, smem, st>>>(x, occ, t2m, n, cin, d, h, w, weight, bias, cout, cout_pad, (__half*)out);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_stem_conv(int dims, const float* x, int n, int cin, int d, int h, int w, const float* weight,
                              const float* bias, int cout, int cout_pad, void* out, void* stream) {
    return stem_launch(dims, x, nullptr, nullptr, n, cin, d, h, w, weight, bias, cout, cout_pad, out, stream);
}

extern "C" int nrrt_stem_conv_maps(int dims, const uint8_t* occ_maps, const int32_t* task_to_map, int n, int cin, int d, int h,
                                   int w, const float* weight, const float* bias, int cout, int cout_pad, void* out,
                                   void* stream) {
    return stem_launch(dims, nullptr, occ_maps, task_to_map, n, cin, d, h, w, weight, bias, cout, cout_pad, out, stream);
}

extern "C" int nrrt_coords_branch(const float* coords, int B, int gh, int gw, const float* const* weights,
                                  const float* const* biases, float* const* feats, void* stream) {
    NRRT_REQUIRE(coords && weights && biases && feats, "null pointer");
    NRRT_REQUIRE(B >= 0 && gh >= 1 && gh <= 3 && gw >= 1 && gw <= 3, "coords grid must be at most 3x3");
    if (B == 0) return NRRT_OK;
    if (B <= 64) {   // too few CTAs for the fused kernel: one launch per layer, parallel over output channels
        const int cins[4] = {1, 64, 128, 256}, couts[4] = {64, 128, 256, 512};
        const float* in = coords;
        for (int l = 0; l < 4; ++l) {
            const size_t smem = (size_t)kTpB * gh * gw * cins[l] * sizeof(float);
            coords_layer_kernel<<<dim3((couts[l] + kCoB - 1) / kCoB, (B * gh * gw + kTpB - 1) / kTpB), kCoB * kTpB, smem,
                                  reinterpret_cast<cudaStream_t>(stream)>>>(in, B, gh, gw, cins[l], couts[l], weights[l], biases[l], feats[l]);
            in = feats[l];
        }
        NRRT_CUDA_TRY(cudaGetLastError());
        return NRRT_OK;
    }
    coords_kernel<<<(B + kTB - 1) / kTB, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(coords, B, gh, gw, weights[0], biases[0], weights[1], biases[1],
                                                                         weights[2], biases[2], weights[3], biases[3], feats[0],
                                                                         feats[1], feats[2], feats[3]);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_coords_fill(const float* feat, int B, int gh, int gw, int C, int dims, int od, int oh, int ow, void* out,
                                int64_t out_cstride, int64_t out_coffset, void* stream) {
    NRRT_REQUIRE(feat && out && C % 8 == 0 && out_cstride % 8 == 0 && out_coffset % 8 == 0, "bad coords_fill arguments");
    const int a0 = dims == 2 ? oh : od, a1 = dims == 2 ? ow : oh, a2 = dims == 2 ? 1 : ow;
    const long long total = (long long)B * a0 * a1 * a2 * (C / 8);
    if (total == 0) return NRRT_OK;
    coords_fill_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(feat, B, gh, gw, C, a0, a1, a2, (__half*)out,
                                                                                                out_cstride, out_coffset);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_coords_poly(const float* feat, int B, int Cc, const float* moments, int cases, int n_uv, int Cout, float a,
                                float b, float* E, void* stream) {
    NRRT_REQUIRE(feat && moments && E, "null pointer");
    NRRT_REQUIRE(B >= 0 && Cc > 0 && Cout > 0 && cases >= 1 && (n_uv == 1 || n_uv == 4), "bad coords_poly shape");
    if (B == 0) return NRRT_OK;
    const size_t smem = (size_t)kPB * 4 * Cc * sizeof(float);
    NRRT_CUDA_TRY(cudaFuncSetAttribute(coords_poly_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    coords_poly_kernel<<<(B + kPB - 1) / kPB, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(feat, B, Cc, moments, cases, n_uv,
                                                                                              Cout, a, b, E);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_gather_channels(const void* src, int64_t src_cstride, int64_t src_coffset, void* dst, int64_t dst_cstride,
                                    int64_t dst_coffset, int C, int64_t pixels, const int32_t* index, int B, void* stream) {
    NRRT_REQUIRE(src && dst && index, "null pointer");
    NRRT_REQUIRE(C % 8 == 0 && src_cstride % 8 == 0 && src_coffset % 8 == 0 && dst_cstride % 8 == 0 && dst_coffset % 8 == 0,
                 "channel counts/offsets must be multiples of 8");
    const long long total = (long long)B * pixels * (C / 8);
    if (total <= 0) return NRRT_OK;
    gather_channels_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        (const __half*)src, src_cstride, src_coffset, (__half*)dst, dst_cstride, dst_coffset, C, pixels, index, B);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_head_1x1(const void* in, int64_t pixels, int C, int64_t in_cstride, const float* weight, float bias,
                             float* logits, void* stream) {
    NRRT_REQUIRE(in && weight && logits && C % 8 == 0 && in_cstride % 8 == 0, "bad head arguments");
    if (pixels == 0) return NRRT_OK;
    head_kernel<<<grid_for(pixels, 256), 256, C * sizeof(float), reinterpret_cast<cudaStream_t>(stream)>>>((const __half*)in, pixels, C,
                                                                                                           in_cstride, weight, bias, logits);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_poly16(const float* e_up, const float* w_taps, const float* e_conv, const float* bias, int B, int Cm,
                           int Cout, int up_h, int up_w, float* work, float* P16, void* stream) {
    NRRT_REQUIRE(e_up && w_taps && e_conv && bias && work && P16, "null pointer");
    NRRT_REQUIRE(B >= 0 && Cm % 16 == 0 && Cout % 128 == 0 && up_h >= 2 && up_w >= 2, "bad poly16 shape");
    if (B == 0) return NRRT_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int M = B * 16, N = 9 * Cout;
    sgemm_kernel<<<dim3(N / GN, (M + GM - 1) / GM), 256, 0, st>>>(e_up, w_taps, work, M, N, Cm);
    NRRT_CUDA_TRY(cudaGetLastError());
    const int H = 2 * up_h, W = 2 * up_w;
    const float a3 = 1.0f / (float)(up_h - 1), b3 = 1.0f / (float)(up_w - 1);
    const float ly = (float)(H - 1) / (2.0f * (float)(up_h - 1)), lx = (float)(W - 1) / (2.0f * (float)(up_w - 1));
    poly16_combine_kernel<<<dim3((Cout + 127) / 128, B), 128, 0, st>>>(work, e_conv, bias, B, Cout, a3, b3, ly, lx, P16);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_poly_relu(const void* pre, const int32_t* pre_index, const float* P16, void* out, int B, int H, int W,
                              int C, void* stream) {
    NRRT_REQUIRE(pre && P16 && out, "null pointer");
    NRRT_REQUIRE(B >= 0 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, "poly_relu needs an even grid (a ConvTranspose k2 s2 output)");
    NRRT_REQUIRE(C % 8 == 0 && C / 8 <= 128 && 256 % (C / 8) == 0, "C/8 must divide 256 (got C = %d)", C);
    NRRT_REQUIRE(((uintptr_t)pre & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)P16 & 15) == 0, "pointers must be 16-byte aligned");
    if (B == 0) return NRRT_OK;
    poly_relu_kernel<<<dim3(H, B), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        (const __half*)pre, pre_index, P16, (__half*)out, H, W, C, 1.0f / (float)(H - 1), 1.0f / (float)(W - 1));
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_coords_pieces_3d(const float* feat, int B, int Cc, float* A, void* stream) {
    NRRT_REQUIRE(feat && A && B >= 0 && Cc > 0, "bad coords_pieces arguments");
    if (B == 0) return NRRT_OK;
    const long long total = (long long)B * 2 * Cc;
    coords_pieces_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(feat, total, Cc, A);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_poly_relu_3d(void* x, const float* E, int B, int D, int H, int W, int C, void* stream) {
    NRRT_REQUIRE(x && E, "null pointer");
    NRRT_REQUIRE(B >= 0 && D >= 2 && H >= 4 && H % 2 == 0 && W >= 2, "poly_relu_3d needs D >= 2, even H >= 4, W >= 2");
    NRRT_REQUIRE(C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0, "C/8 must divide 256 (got C = %d)", C);
    NRRT_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)E & 15) == 0, "pointers must be 16-byte aligned");
    if (B == 0) return NRRT_OK;
    poly_relu3d_kernel<<<dim3(D * H, B), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>((__half*)x, E, D, H, W, C,
                                                                                         1.0f / (float)(D - 1), 1.0f / (float)(H - 1));
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_coords_poly_gemm(const float* feat, int B, int Cc, const float* moments_t, int cases, int n_uv, int Cout,
                                     float a, float b, float* work, float* E, void* stream) {
    NRRT_REQUIRE(feat && moments_t && work && E, "null pointer");
    NRRT_REQUIRE(B >= 0 && Cc % 16 == 0 && Cout > 0 && cases >= 1 && (n_uv == 1 || n_uv == 4) && (cases * n_uv * Cout) % GN == 0,
                 "bad coords_poly_gemm shape");
    if (B == 0) return NRRT_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int N = cases * n_uv * Cout, M = B * 4;
    const bool pre_g = cases == 54;         // 3D: feat already holds the polynomial coefficients A (nrrt_coords_pieces_3d)
    const float* G = pre_g ? feat : work;   // [B][4][Cc]
    float* Cw = work + (size_t)B * 4 * Cc;  // [B*4][N]
    if (!pre_g) {
        coords_g_kernel<<<grid_for((long long)B * Cc, 256), 256, 0, st>>>(feat, (long long)B * Cc, Cc, work);
        NRRT_CUDA_TRY(cudaGetLastError());
    }
    sgemm_kernel<<<dim3(N / GN, (M + GM - 1) / GM), 256, 0, st>>>(G, moments_t, Cw, M, N, Cc);
    NRRT_CUDA_TRY(cudaGetLastError());
    const long long total = (long long)B * cases * Cout;
    coords_combine_kernel<<<grid_for(total, 256), 256, 0, st>>>(Cw, total, cases, n_uv, Cout, a, b, E);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_poly_fold(const float* E, const float* cases_const, int B, int Cout, int H, int W, float* Ef, void* stream) {
    NRRT_REQUIRE(E && Ef && B >= 0 && Cout > 0 && H >= 2 && W >= 2, "bad poly_fold arguments");
    if (B == 0) return NRRT_OK;
    const long long total = (long long)B * 36 * Cout;
    const float inv_oh = 1.0f / (float)(2 * H - 1), inv_ow = 1.0f / (float)(2 * W - 1);
    poly_fold_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        E, cases_const, B, Cout, 2.0f * (float)(H - 1) * inv_oh, 2.0f * (float)(W - 1) * inv_ow, inv_oh, inv_ow, Ef);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}
