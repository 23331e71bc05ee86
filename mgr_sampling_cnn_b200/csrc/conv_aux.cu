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
                            if (omap) v = __ldg(omap + ((long long)zz * H + yy) * W + xx) != 0 ? 1.f : 0.f;
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

int grid_for(long long work, int block) {
    long long g = (work + block - 1) / block;
    return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
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
