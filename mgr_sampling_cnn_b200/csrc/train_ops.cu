// train_ops.cu — the HBM-bound kernels of the training step (UNet_cooler.training_step train.py:205-210 with
// nn.BCEWithLogitsLoss :113, torch.optim.Adam(lr=1e-3) :228-229, train-mode BatchNorm2d of the ResNet blocks).
// The contractions (forward convs, data gradients = the same tcgen05 kernels with transposed weights, weight gradients =
// wgrad_tc.cu) run on the tensor cores; everything here is one pass over an activation tensor: channels-last fp16 in,
// fp32 arithmetic, fp16 out, per-channel reductions accumulated per block and added with one atomic per channel.
#include <cuda_fp16.h>

#include "nrrt_common.cuh"

namespace {

constexpr int kEw = 256;   // threads per block of the channels-last kernels

// A thread owns one 8-channel chunk of the pixels p = blockIdx.x * ppb + lane_p, + gridDim.x * ppb, ...
struct Walk {
    int chunks, ppb, chunk, lane_p;
    bool active;
    __device__ Walk(int C) {
        chunks = C >> 3;
        ppb = kEw / chunks;
        if (ppb < 1) ppb = 1;
        chunk = threadIdx.x % chunks;
        lane_p = threadIdx.x / chunks;
        active = lane_p < ppb && chunks <= kEw;
    }
};

__device__ __forceinline__ void load8(const __half* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        v[2 * j] = f.x; v[2 * j + 1] = f.y;
    }
}
__device__ __forceinline__ void store8(__half* p, const float (&v)[8]) {
    uint4 u;
    __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    *reinterpret_cast<uint4*>(p) = u;
}

// block reduction of K per-channel accumulators (8 channels per thread) -> atomics on global fp64 / fp32 arrays
template <int K, typename T>
__device__ __forceinline__ void reduce_channels(const Walk& w, float (&acc)[K][8], T* const (&dst)[K], int C) {
    extern __shared__ float red[];   // [ppb][C]
    for (int k = 0; k < K; ++k) {
        __syncthreads();
        if (w.active)
#pragma unroll
            for (int j = 0; j < 8; ++j) red[w.lane_p * C + w.chunk * 8 + j] = acc[k][j];
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += kEw) {
            float s = 0.f;
            for (int q = 0; q < w.ppb; ++q) s += red[q * C + c];
            if (dst[k]) atomicAdd(dst[k] + c, (T)s);
        }
    }
}

// g = dy where y > 0 else 0 (y = the layer's stored ReLU output); dbias[c] += sum g          (conv + bias + ReLU layers)
__global__ void __launch_bounds__(kEw) relu_bwd_sum_kernel(const __half* dy, long long dy_cs, const __half* y, long long y_cs, __half* g,
                                                           long long g_cs, float* dbias, long long M, int C) {
    const Walk w(C);
    float acc[1][8] = {};
    if (w.active)
        for (long long p = (long long)blockIdx.x * w.ppb + w.lane_p; p < M; p += (long long)gridDim.x * w.ppb) {
            float a[8], m[8];
            load8(dy + p * dy_cs + w.chunk * 8, a);
            if (y) {
                load8(y + p * y_cs + w.chunk * 8, m);
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = m[j] > 0.f ? a[j] : 0.f;
            }
            if (g) store8(g + p * g_cs + w.chunk * 8, a);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[0][j] += a[j];
        }
    float* const dst[1] = {dbias};
    if (dbias) reduce_channels<1, float>(w, acc, dst, C);
}

// train-mode BatchNorm statistics: sum[c] += sum z, sumsq[c] += sum z^2 (fp64 accumulators)
__global__ void __launch_bounds__(kEw) bn_stats_kernel(const __half* z, long long z_cs, double* sum, double* sumsq, long long M, int C) {
    const Walk w(C);
    float acc[2][8] = {};
    if (w.active) {
        const long long stride = (long long)gridDim.x * w.ppb;
        long long p = (long long)blockIdx.x * w.ppb + w.lane_p;
        for (; p + 3 * stride < M; p += 4 * stride) {          // four independent 16-byte loads in flight per thread
            float a[4][8];
#pragma unroll
            for (int u = 0; u < 4; ++u) load8(z + (p + u * stride) * z_cs + w.chunk * 8, a[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int j = 0; j < 8; ++j) { acc[0][j] += a[u][j]; acc[1][j] += a[u][j] * a[u][j]; }
        }
        for (; p < M; p += stride) {
            float a[8];
            load8(z + p * z_cs + w.chunk * 8, a);
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc[0][j] += a[j]; acc[1][j] += a[j] * a[j]; }
        }
    }
    double* const dst[2] = {sum, sumsq};
    reduce_channels<2, double>(w, acc, dst, C);
}

// mean / biased variance -> (scale, shift) of the normalisation, saved (mean, rstd), running statistics (momentum, unbiased)
__global__ void bn_finalize_kernel(const double* sum, const double* sumsq, long long M, const float* gamma, const float* beta, float eps,
                                   float momentum, float* running_mean, float* running_var, float* mean, float* rstd, float* scale,
                                   float* shift, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double mu = sum[c] / (double)M;
    double var = sumsq[c] / (double)M - mu * mu;
    if (var < 0) var = 0;
    const float rs = (float)(1.0 / sqrt(var + (double)eps));
    mean[c] = (float)mu;
    rstd[c] = rs;
    scale[c] = gamma[c] * rs;
    shift[c] = beta[c] - (float)mu * gamma[c] * rs;
    if (running_mean) {
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
        const double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
}

// y = [relu](z * scale[c] + shift[c] [+ res])
__global__ void __launch_bounds__(kEw) affine_act_kernel(const __half* z, long long z_cs, const float* scale, const float* shift,
                                                         const __half* res, long long res_cs, int relu, __half* y, long long y_cs,
                                                         long long M, int C) {
    const Walk w(C);
    if (!w.active) return;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = scale[w.chunk * 8 + j]; sh[j] = shift[w.chunk * 8 + j]; }
    for (long long p = (long long)blockIdx.x * w.ppb + w.lane_p; p < M; p += (long long)gridDim.x * w.ppb) {
        float a[8], r[8];
        load8(z + p * z_cs + w.chunk * 8, a);
        if (res) load8(res + p * res_cs + w.chunk * 8, r);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v = a[j] * sc[j] + sh[j];
            if (res) v += r[j];
            a[j] = relu ? fmaxf(v, 0.f) : v;
        }
        store8(y + p * y_cs + w.chunk * 8, a);
    }
}

// BatchNorm backward, pass 1: g = dy masked by y > 0 (optional); dbeta[c] += sum g; dgamma[c] += sum g * xhat
__global__ void __launch_bounds__(kEw) bn_bwd_reduce_kernel(const __half* dy, long long dy_cs, const __half* y, long long y_cs,
                                                            const __half* z, long long z_cs, const float* mean, const float* rstd,
                                                            double* dbeta, double* dgamma, long long M, int C) {
    const Walk w(C);
    float acc[2][8] = {};
    if (w.active) {
        float mu[8], rs[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { mu[j] = mean[w.chunk * 8 + j]; rs[j] = rstd[w.chunk * 8 + j]; }
        for (long long p = (long long)blockIdx.x * w.ppb + w.lane_p; p < M; p += (long long)gridDim.x * w.ppb) {
            float a[8], m[8], zz[8];
            load8(dy + p * dy_cs + w.chunk * 8, a);
            load8(z + p * z_cs + w.chunk * 8, zz);
            if (y) {
                load8(y + p * y_cs + w.chunk * 8, m);
#pragma unroll
                for (int j = 0; j < 8; ++j) a[j] = m[j] > 0.f ? a[j] : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc[0][j] += a[j]; acc[1][j] += a[j] * (zz[j] - mu[j]) * rs[j]; }
        }
    }
    double* const dst[2] = {dbeta, dgamma};
    reduce_channels<2, double>(w, acc, dst, C);
}

// pass 2: dz = gamma * rstd * (g - dbeta / M - xhat * dgamma / M); optionally g_out = g (the residual branch's gradient)
__global__ void __launch_bounds__(kEw) bn_bwd_apply_kernel(const __half* dy, long long dy_cs, const __half* y, long long y_cs,
                                                           const __half* z, long long z_cs, const float* mean, const float* rstd,
                                                           const float* gamma, const double* dbeta, const double* dgamma, __half* dz,
                                                           long long dz_cs, __half* g_out, long long g_cs, long long M, int C) {
    const Walk w(C);
    if (!w.active) return;
    float mu[8], rs[8], k0[8], k1[8], k2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = w.chunk * 8 + j;
        mu[j] = mean[c]; rs[j] = rstd[c];
        k0[j] = gamma[c] * rs[j];
        k1[j] = (float)(dbeta[c] / (double)M);
        k2[j] = (float)(dgamma[c] / (double)M);
    }
    for (long long p = (long long)blockIdx.x * w.ppb + w.lane_p; p < M; p += (long long)gridDim.x * w.ppb) {
        float a[8], m[8], zz[8];
        load8(dy + p * dy_cs + w.chunk * 8, a);
        load8(z + p * z_cs + w.chunk * 8, zz);
        if (y) {
            load8(y + p * y_cs + w.chunk * 8, m);
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = m[j] > 0.f ? a[j] : 0.f;
        }
        if (g_out) store8(g_out + p * g_cs + w.chunk * 8, a);
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = k0[j] * (a[j] - k1[j] - (zz[j] - mu[j]) * rs[j] * k2[j]);
        store8(dz + p * dz_cs + w.chunk * 8, a);
    }
}

// channels-last pixel remaps (fp16, 8-channel chunks): mode 0 = space to depth (in [N][2H][2W][C] -> out [N][H][W][4C],
// channel (a * 2 + b) * C + c = in pixel (2h + a, 2w + b)); mode 1 = zero-stuffing x2 (in [N][H][W][C] -> out [N][2H][2W][C],
// value at even positions, out pre-zeroed)
__global__ void __launch_bounds__(kEw) remap_kernel(const __half* in, long long in_cs, __half* out, long long out_cs, int N, int H, int W,
                                                    int C, int mode) {
    const int chunks = C >> 3;
    const long long total = (long long)N * H * W * chunks * (mode == 0 ? 4 : 1);
    for (long long i = (long long)blockIdx.x * kEw + threadIdx.x; i < total; i += (long long)gridDim.x * kEw) {
        const int ch = (int)(i % chunks);
        long long t = i / chunks;
        if (mode == 0) {
            const int g = (int)(t & 3); t >>= 2;
            const int w = (int)(t % W); t /= W;
            const int h = (int)(t % H);
            const int n = (int)(t / H);
            const size_t src = (((size_t)n * 2 * H + 2 * h + (g >> 1)) * 2 * W + 2 * w + (g & 1)) * in_cs + ch * 8;
            const size_t dst = (((size_t)n * H + h) * W + w) * out_cs + (size_t)g * C + ch * 8;
            *reinterpret_cast<uint4*>(out + dst) = *reinterpret_cast<const uint4*>(in + src);
        } else {
            const int w = (int)(t % W); t /= W;
            const int h = (int)(t % H);
            const int n = (int)(t / H);
            const size_t src = (((size_t)n * H + h) * W + w) * in_cs + ch * 8;
            const size_t dst = (((size_t)n * 2 * H + 2 * h) * 2 * W + 2 * w) * out_cs + ch * 8;
            *reinterpret_cast<uint4*>(out + dst) = *reinterpret_cast<const uint4*>(in + src);
        }
    }
}

// nn.BCEWithLogitsLoss (mean): loss_sum += sum max(z, 0) - z t + log1p(exp(-|z|)); dlogits = (sigmoid(z) - t) * grad_scale
__global__ void __launch_bounds__(256) bce_kernel(float* logits, const float* logit_bias, const float* target, float* dlogits,
                                                  double* loss_sum, float grad_scale, const float* scale_state, long long M) {
    double local = 0.0;
    if (scale_state) grad_scale = scale_state[0];
    const float zb = logit_bias ? *logit_bias : 0.f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < M; i += (long long)gridDim.x * 256) {
        const float z = logits[i] + zb, t = target[i];
        if (logit_bias) logits[i] = z;
        const float e = expf(-fabsf(z));
        local += (double)(fmaxf(z, 0.f) - z * t + log1pf(e));
        const float sig = z >= 0.f ? 1.f / (1.f + e) : e / (1.f + e);
        if (dlogits) dlogits[i] = (sig - t) * grad_scale;
    }
    __shared__ double red[256];
    red[threadIdx.x] = local;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(loss_sum, red[0]);
}

// MapsDataset's mask normalisation (train.py:49-50): t = (m - min) / (max - min) in float64, stored as float32; one block per sample
__global__ void __launch_bounds__(256) mask_targets_kernel(const uint8_t* masks, float* target, long long cells) {
    const uint8_t* m = masks + (size_t)blockIdx.x * cells;
    float* t = target + (size_t)blockIdx.x * cells;
    __shared__ int smin[256], smax[256];
    int lo = 255, hi = 0;
    for (long long i = threadIdx.x; i < cells; i += 256) { const int v = m[i]; lo = min(lo, v); hi = max(hi, v); }
    smin[threadIdx.x] = lo; smax[threadIdx.x] = hi;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) { smin[threadIdx.x] = min(smin[threadIdx.x], smin[threadIdx.x + s]); smax[threadIdx.x] = max(smax[threadIdx.x], smax[threadIdx.x + s]); }
        __syncthreads();
    }
    lo = smin[0]; hi = smax[0];
    const double den = (double)(hi - lo);   // 0 -> NaN / inf exactly as NumPy's division does
    for (long long i = threadIdx.x; i < cells; i += 256) t[i] = (float)((double)(m[i] - lo) / den);
}

// final_conv backward (1x1, 64 -> 1; train.py:168,202): g8 = x8 > 0 ? dl * w[c] : 0; dw[c] += sum dl * x8; db += sum dl
__global__ void __launch_bounds__(kEw) head_bwd_kernel(const float* dl, const __half* x8, long long x_cs, const float* w, __half* g8,
                                                       long long g_cs, float* dw, float* db, long long M, int C) {
    const Walk wk(C);
    float acc[1][8] = {};
    float dbl = 0.f;
    if (wk.active) {
        float wc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) wc[j] = w[wk.chunk * 8 + j];
        for (long long p = (long long)blockIdx.x * wk.ppb + wk.lane_p; p < M; p += (long long)gridDim.x * wk.ppb) {
            float x[8], g[8];
            const float d = dl[p];
            load8(x8 + p * x_cs + wk.chunk * 8, x);
#pragma unroll
            for (int j = 0; j < 8; ++j) { g[j] = x[j] > 0.f ? d * wc[j] : 0.f; acc[0][j] += d * x[j]; }
            store8(g8 + p * g_cs + wk.chunk * 8, g);
            if (wk.chunk == 0) dbl += d;
        }
    }
    float* const dst[1] = {dw};
    reduce_channels<1, float>(wk, acc, dst, C);
    __shared__ float sdb[kEw];
    __syncthreads();
    sdb[threadIdx.x] = dbl;
    __syncthreads();
    for (int s = kEw / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) sdb[threadIdx.x] += sdb[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(db, sdb[0]);
}

// torch.optim.Adam (no weight decay, no amsgrad): one flat fp32 parameter buffer
__global__ void __launch_bounds__(256) adam_kernel(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2,
                                                   float eps, float bc1, float bc2_sqrt, float grad_scale, const float* scale_state) {
    if (scale_state) {   // dynamic loss scaling: [0] = the scale the gradients carry, [1] = good steps, [2] = overflow flag, [3] = steps taken
        if (scale_state[2] != 0.f) return;                       // a non-finite gradient: the step is skipped (nrrt_loss_scale_update backs off)
        grad_scale /= scale_state[0];
        const float t = scale_state[3] + 1.f;
        bc1 = 1.f - powf(b1, t);
        bc2_sqrt = sqrtf(1.f - powf(b2, t));
    }
    const float step = lr / bc1;
    auto upd = [&](float& pi, float gi, float& mi, float& vi) {
        gi *= grad_scale;
        mi = b1 * mi + (1.f - b1) * gi;
        vi = b2 * vi + (1.f - b2) * gi * gi;
        pi -= step * (mi / (sqrtf(vi) / bc2_sqrt + eps));
    };
    const long long n4 = n >> 2;                                 // 16-byte accesses (the buffers are 256-byte aligned torch allocations)
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    const bool aligned = ((((uintptr_t)p) | ((uintptr_t)g) | ((uintptr_t)m) | ((uintptr_t)v)) & 15) == 0;
    if (aligned) {
        for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
            float4 pp = p4[i], mm = m4[i], vv = v4[i];
            const float4 gg = g4[i];
            upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
            p4[i] = pp; m4[i] = mm; v4[i] = vv;
        }
    }
    for (long long i = (aligned ? n4 * 4 : 0) + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256)
        upd(p[i], g[i], m[i], v[i]);
}

__global__ void __launch_bounds__(256) to_half_kernel(const float* src, __half* dst, long long n) {
    const bool aligned = ((((uintptr_t)src) | ((uintptr_t)dst)) & 15) == 0;
    const long long n8 = aligned ? n >> 3 : 0;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n8; i += (long long)gridDim.x * 256) {
        const float4 a = reinterpret_cast<const float4*>(src)[2 * i], b = reinterpret_cast<const float4*>(src)[2 * i + 1];
        uint4 o;
        __half2* h = reinterpret_cast<__half2*>(&o);
        h[0] = __floats2half2_rn(a.x, a.y); h[1] = __floats2half2_rn(a.z, a.w);
        h[2] = __floats2half2_rn(b.x, b.y); h[3] = __floats2half2_rn(b.z, b.w);
        reinterpret_cast<uint4*>(dst)[i] = o;
    }
    for (long long i = n8 * 8 + (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) dst[i] = __float2half_rn(src[i]);
}

// data-gradient weights of a conv: dst fp16 [ci_rows][taps][co_cols] with dst[ci][taps-1-t][co] = src[co][t][ci]
// (src fp32 [co][taps][ci_stride]; rows ci >= cin_valid and columns co >= co are zero)
__global__ void __launch_bounds__(256) pack_dgrad_kernel(const float* src, __half* dst, int co, int taps, int ci_stride, int ci_rows,
                                                         int co_cols) {
    const long long total = (long long)ci_rows * taps * co_cols;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int o = (int)(i % co_cols);
        const int t = (int)((i / co_cols) % taps);
        const int c = (int)(i / ((long long)co_cols * taps));
        float v = 0.f;
        if (o < co && c < ci_stride) v = src[((size_t)o * taps + (taps - 1 - t)) * ci_stride + c];
        dst[i] = __float2half_rn(v);
    }
}

// the same for every convolution of a model in ONE launch: blockIdx.y = layer, desc = 8 int64 per layer
// {src offset (floats), dst offset (halfs), co, taps, ci_stride, ci_rows, co_cols, 0}
__global__ void __launch_bounds__(256) pack_dgrad_batched_kernel(const float* params, __half* dst_base, const long long* descs) {
    const long long* d = descs + 8 * blockIdx.y;
    const float* src = params + d[0];
    __half* dst = dst_base + d[1];
    const int co = (int)d[2], taps = (int)d[3], ci_stride = (int)d[4], ci_rows = (int)d[5], co_cols = (int)d[6];
    const long long total = (long long)ci_rows * taps * co_cols;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int o = (int)(i % co_cols);
        const int t = (int)((i / co_cols) % taps);
        const int c = (int)(i / ((long long)co_cols * taps));
        float v = 0.f;
        if (o < co && c < ci_stride) v = src[((size_t)o * taps + (taps - 1 - t)) * ci_stride + c];
        dst[i] = __float2half_rn(v);
    }
}

// first conv (3 -> 32, train.py:118) weight gradient from the fp32 NCHW input: dW[co][ci][tap] += sum g[p][co] x[ci][p + tap]
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* x, const __half* g, long long g_cs, int N, int cin, int H, int W, int cout,
                                                         float* dw, float* db) {
    constexpr int T = 16;
    extern __shared__ float sm[];
    float* xs = sm;                               // [cin][T + 2][T + 2]
    float* gs = sm + cin * (T + 2) * (T + 2);     // [T * T][cout]
    const int tiles_w = (W + T - 1) / T, tiles_h = (H + T - 1) / T;
    int t = blockIdx.x;
    const int tw = t % tiles_w; t /= tiles_w;
    const int th = t % tiles_h;
    const int n = t / tiles_h;
    const int h0 = th * T, w0 = tw * T;
    for (int i = threadIdx.x; i < cin * (T + 2) * (T + 2); i += 256) {
        const int xx = i % (T + 2), yy = (i / (T + 2)) % (T + 2), c = i / ((T + 2) * (T + 2));
        const int h = h0 + yy - 1, w = w0 + xx - 1;
        xs[i] = (h >= 0 && h < H && w >= 0 && w < W) ? x[(((size_t)n * cin + c) * H + h) * W + w] : 0.f;
    }
    for (int i = threadIdx.x; i < T * T * cout; i += 256) {
        const int c = i % cout, px = i / cout;
        const int h = h0 + px / T, w = w0 + px % T;
        gs[i] = (h < H && w < W) ? __half2float(g[(((size_t)n * H + h) * W + w) * g_cs + c]) : 0.f;
    }
    __syncthreads();
    const int outs = cout * cin * 9;
    for (int o = threadIdx.x; o < outs; o += 256) {
        const int tap = o % 9, c = (o / 9) % cin, k = o / (9 * cin);
        const int ky = tap / 3, kx = tap % 3;
        float s = 0.f;
        for (int py = 0; py < T; ++py)
            for (int px = 0; px < T; ++px) s += gs[(py * T + px) * cout + k] * xs[(c * (T + 2) + py + ky) * (T + 2) + px + kx];
        atomicAdd(dw + o, s);
    }
    if (db)
        for (int k = threadIdx.x; k < cout; k += 256) {
            float s = 0.f;
            for (int px = 0; px < T * T; ++px) s += gs[px * cout + k];
            atomicAdd(db + k, s);
        }
}

// ---- coordinate branch backward (train.py:179-190) -----------------------------------------------------------------------
// dfeat[b][q][c] += sum over pixels of bilinear weight_q(pixel) * dcat[b][pixel][coff + c]  (F.interpolate align_corners=True, 2x2 source)
__global__ void __launch_bounds__(kEw) coords_fill_bwd_kernel(const __half* dcat, long long cs, int H, int W, int C, float* dfeat) {
    const Walk w(C);
    const int b = blockIdx.y;
    float acc[4][8] = {};
    const long long M = (long long)H * W;
    const float ih = H > 1 ? 1.f / (float)(H - 1) : 0.f, iw = W > 1 ? 1.f / (float)(W - 1) : 0.f;
    if (w.active)
        for (long long p = (long long)blockIdx.x * w.ppb + w.lane_p; p < M; p += (long long)gridDim.x * w.ppb) {
            float a[8];
            load8(dcat + ((size_t)b * M + p) * cs + w.chunk * 8, a);
            const float ty = (float)(p / W) * ih, tx = (float)(p % W) * iw;
            const float w00 = (1.f - ty) * (1.f - tx), w01 = (1.f - ty) * tx, w10 = ty * (1.f - tx), w11 = ty * tx;
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc[0][j] += w00 * a[j]; acc[1][j] += w01 * a[j]; acc[2][j] += w10 * a[j]; acc[3][j] += w11 * a[j]; }
        }
    float* const dst[4] = {dfeat + ((size_t)b * 4 + 0) * C, dfeat + ((size_t)b * 4 + 1) * C, dfeat + ((size_t)b * 4 + 2) * C,
                           dfeat + ((size_t)b * 4 + 3) * C};
    reduce_channels<4, float>(w, acc, dst, C);
}

// one conv_coords layer (3x3, pad 1, on the 2x2 grid, + ReLU): g = dfeat_out masked by f_out > 0;
// dW[tap][ci][co] += sum_b,p,q g[b][p][co] f_in[b][q][ci] (tap = (qy - py + 1) * 3 + qx - px + 1); db[co] += sum g
__global__ void __launch_bounds__(256) coords_wgrad_kernel(const float* dfeat_out, const float* f_out, const float* f_in, int B, int Cin,
                                                           int Cout, float* dW, float* db) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long long)Cin * Cout) return;
    const int co = (int)(i % Cout), ci = (int)(i / Cout);
    float acc[9] = {};
    float bsum = 0.f;
    for (int b = 0; b < B; ++b) {
        float g[4], f[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const size_t o = ((size_t)b * 4 + p) * Cout + co;
            g[p] = f_out[o] > 0.f ? dfeat_out[o] : 0.f;
            f[p] = f_in[((size_t)b * 4 + p) * Cin + ci];
            bsum += g[p];
        }
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[((q >> 1) - (p >> 1) + 1) * 3 + (q & 1) - (p & 1) + 1] += g[p] * f[q];
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) dW[((size_t)t * Cin + ci) * Cout + co] += acc[t];
    if (ci == 0) db[co] += bsum;
}

// dfeat_in[b][q][ci] += sum_p,co W[tap(p,q)][ci][co] g[b][p][co]
__global__ void __launch_bounds__(128) coords_dgrad_kernel(const float* dfeat_out, const float* f_out, const float* Wt, int B, int Cin, int Cout,
                                                           float* dfeat_in) {
    // one block per (b, q, ci); threads stride over co
    int t = blockIdx.x;
    const int ci = t % Cin; t /= Cin;
    const int q = t & 3;
    const int b = t >> 2;
    float s = 0.f;
    for (int co = threadIdx.x; co < Cout; co += 128)
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const size_t o = ((size_t)b * 4 + p) * Cout + co;
            const float g = f_out[o] > 0.f ? dfeat_out[o] : 0.f;
            const int tap = ((q >> 1) - (p >> 1) + 1) * 3 + (q & 1) - (p & 1) + 1;
            s += Wt[((size_t)tap * Cin + ci) * Cout + co] * g;
        }
    __shared__ float red[128];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int k = 64; k > 0; k >>= 1) {
        if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) dfeat_in[((size_t)b * 4 + q) * Cin + ci] += red[0];
}

int ew_grid(long long M, int C) {
    const int chunks = C / 8;
    const int ppb = kEw / chunks > 0 ? kEw / chunks : 1;
    long long blocks = (M + ppb - 1) / ppb;
    const long long cap = 148 * 8;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}
int ew_grid_red(long long M, int C) {   // reduction kernels: fewer, longer blocks (every block ends with C atomics per accumulator)
    const int g = ew_grid(M, C);
    return g < 148 * 3 ? g : 148 * 3;
}
size_t ew_smem(int C) {
    const int chunks = C / 8;
    const int ppb = kEw / chunks > 0 ? kEw / chunks : 1;
    return (size_t)ppb * C * sizeof(float);
}

#define EW_CHECK(C)                                                                                                              \
    NRRT_REQUIRE((C) >= 8 && (C) % 8 == 0 && (C) <= 2048, "channels must be a multiple of 8, at most 2048 (got %d)", (int)(C))

}  // namespace

extern "C" int nrrt_relu_bwd_sum(const void* dy, int64_t dy_cs, const void* y, int64_t y_cs, void* g, int64_t g_cs, float* dbias,
                                 int64_t pixels, int c, void* stream) {
    EW_CHECK(c);
    NRRT_REQUIRE(dy && pixels >= 1 && dy_cs % 8 == 0 && y_cs % 8 == 0 && g_cs % 8 == 0, "bad arguments");
    relu_bwd_sum_kernel<<<ew_grid(pixels, c), kEw, ew_smem(c), (cudaStream_t)stream>>>((const __half*)dy, dy_cs, (const __half*)y, y_cs,
                                                                                       (__half*)g, g_cs, dbias, pixels, c);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_bn_train_forward(const void* z, int64_t z_cs, int64_t pixels, int c, const float* gamma, const float* beta, float eps,
                                     float momentum, float* running_mean, float* running_var, const void* res, int64_t res_cs, int relu,
                                     void* y, int64_t y_cs, float* save_mean, float* save_rstd, float* scale_shift, double* work,
                                     void* stream) {
    EW_CHECK(c);
    NRRT_REQUIRE(z && y && gamma && beta && save_mean && save_rstd && scale_shift && work && pixels >= 1, "null pointer");
    NRRT_REQUIRE(z_cs % 8 == 0 && y_cs % 8 == 0 && res_cs % 8 == 0, "channel strides must be multiples of 8");
    cudaStream_t st = (cudaStream_t)stream;
    NRRT_CUDA_TRY(cudaMemsetAsync(work, 0, 2 * (size_t)c * sizeof(double), st));
    bn_stats_kernel<<<ew_grid_red(pixels, c), kEw, ew_smem(c), st>>>((const __half*)z, z_cs, work, work + c, pixels, c);
    bn_finalize_kernel<<<(c + 127) / 128, 128, 0, st>>>(work, work + c, pixels, gamma, beta, eps, momentum, running_mean, running_var,
                                                       save_mean, save_rstd, scale_shift, scale_shift + c, c);
    affine_act_kernel<<<ew_grid(pixels, c), kEw, 0, st>>>((const __half*)z, z_cs, scale_shift, scale_shift + c, (const __half*)res, res_cs,
                                                          relu, (__half*)y, y_cs, pixels, c);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

namespace {
__global__ void bn_eval_params_kernel(const float* gamma, const float* beta, const float* running_mean, const float* running_var, float eps,
                                      float* scale, float* shift, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float sc = gamma[c] / sqrtf(running_var[c] + eps);
    scale[c] = sc;
    shift[c] = beta[c] - running_mean[c] * sc;
}
}  // namespace

extern "C" int nrrt_bn_eval_forward(const void* z, int64_t z_cs, int64_t pixels, int c, const float* gamma, const float* beta, float eps,
                                    const float* running_mean, const float* running_var, const void* res, int64_t res_cs, int relu, void* y,
                                    int64_t y_cs, float* scale_shift, void* stream) {
    EW_CHECK(c);
    NRRT_REQUIRE(z && y && gamma && beta && running_mean && running_var && scale_shift && pixels >= 1, "null pointer");
    NRRT_REQUIRE(z_cs % 8 == 0 && y_cs % 8 == 0 && res_cs % 8 == 0, "channel strides must be multiples of 8");
    cudaStream_t st = (cudaStream_t)stream;
    bn_eval_params_kernel<<<(c + 127) / 128, 128, 0, st>>>(gamma, beta, running_mean, running_var, eps, scale_shift, scale_shift + c, c);
    affine_act_kernel<<<ew_grid(pixels, c), kEw, 0, st>>>((const __half*)z, z_cs, scale_shift, scale_shift + c, (const __half*)res, res_cs,
                                                          relu, (__half*)y, y_cs, pixels, c);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_bn_train_backward(const void* dy, int64_t dy_cs, const void* y_mask, int64_t y_cs, const void* z, int64_t z_cs,
                                      int64_t pixels, int c, const float* gamma, const float* save_mean, const float* save_rstd,
                                      void* dz, int64_t dz_cs, void* g_out, int64_t g_cs, float* dgamma, float* dbeta, double* work,
                                      void* stream);

namespace {
__global__ void bn_bwd_params_kernel(const double* work, float* dgamma, float* dbeta, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    dbeta[c] += (float)work[c];
    dgamma[c] += (float)work[C + c];
}
}  // namespace

extern "C" int nrrt_bn_train_backward(const void* dy, int64_t dy_cs, const void* y_mask, int64_t y_cs, const void* z, int64_t z_cs,
                                      int64_t pixels, int c, const float* gamma, const float* save_mean, const float* save_rstd,
                                      void* dz, int64_t dz_cs, void* g_out, int64_t g_cs, float* dgamma, float* dbeta, double* work,
                                      void* stream) {
    EW_CHECK(c);
    NRRT_REQUIRE(dy && z && dz && gamma && save_mean && save_rstd && dgamma && dbeta && work && pixels >= 1, "null pointer");
    NRRT_REQUIRE(dy_cs % 8 == 0 && y_cs % 8 == 0 && z_cs % 8 == 0 && dz_cs % 8 == 0 && g_cs % 8 == 0, "channel strides must be multiples of 8");
    cudaStream_t st = (cudaStream_t)stream;
    NRRT_CUDA_TRY(cudaMemsetAsync(work, 0, 2 * (size_t)c * sizeof(double), st));
    bn_bwd_reduce_kernel<<<ew_grid_red(pixels, c), kEw, ew_smem(c), st>>>((const __half*)dy, dy_cs, (const __half*)y_mask, y_cs, (const __half*)z,
                                                                      z_cs, save_mean, save_rstd, work, work + c, pixels, c);
    bn_bwd_apply_kernel<<<ew_grid(pixels, c), kEw, 0, st>>>((const __half*)dy, dy_cs, (const __half*)y_mask, y_cs, (const __half*)z, z_cs,
                                                            save_mean, save_rstd, gamma, work, work + c, (__half*)dz, dz_cs, (__half*)g_out,
                                                            g_cs, pixels, c);
    bn_bwd_params_kernel<<<(c + 127) / 128, 128, 0, st>>>(work, dgamma, dbeta, c);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_remap_channels_last(const void* in, int64_t in_cs, void* out, int64_t out_cs, int n, int h, int w, int c, int mode,
                                        void* stream) {
    NRRT_REQUIRE(in && out && n >= 1 && h >= 1 && w >= 1 && c >= 8 && c % 8 == 0 && in_cs % 8 == 0 && out_cs % 8 == 0, "bad arguments");
    NRRT_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (space to depth; h, w = OUTPUT grid) or 1 (zero-stuffing; h, w = INPUT grid)");
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 1) {
        NRRT_REQUIRE(out_cs == c, "zero-stuffing writes a dense output (out_cstride == c)");
        NRRT_CUDA_TRY(cudaMemsetAsync(out, 0, (size_t)n * h * w * 4 * c * 2, st));
    }
    const long long total = (long long)n * h * w * (c / 8) * (mode == 0 ? 4 : 1);
    const long long blocks = (total + kEw - 1) / kEw;
    remap_kernel<<<(int)(blocks < 148 * 16 ? blocks : 148 * 16), kEw, 0, st>>>((const __half*)in, in_cs, (__half*)out, out_cs, n, h, w, c, mode);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_bce_with_logits(float* logits, const float* logit_bias, const float* target, int64_t count, float grad_scale,
                                    const float* scale_state, float* dlogits, double* loss_sum, void* stream) {
    NRRT_REQUIRE(logits && target && loss_sum && count >= 1, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    NRRT_CUDA_TRY(cudaMemsetAsync(loss_sum, 0, sizeof(double), st));
    const long long blocks = (count + 255) / 256;
    bce_kernel<<<(int)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, st>>>(logits, logit_bias, target, dlogits, loss_sum, grad_scale, scale_state, count);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_mask_targets(const uint8_t* masks, int B, int64_t cells, float* target, void* stream) {
    NRRT_REQUIRE(masks && target && B >= 1 && cells >= 1, "bad arguments");
    mask_targets_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(masks, target, cells);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_head_backward(const float* dlogits, const void* x, int64_t x_cs, const float* weight, int64_t pixels, int c, void* g,
                                  int64_t g_cs, float* dweight, float* dbias, void* stream) {
    EW_CHECK(c);
    NRRT_REQUIRE(dlogits && x && weight && g && dweight && dbias && pixels >= 1 && x_cs % 8 == 0 && g_cs % 8 == 0, "bad arguments");
    head_bwd_kernel<<<ew_grid(pixels, c), kEw, ew_smem(c), (cudaStream_t)stream>>>(dlogits, (const __half*)x, x_cs, weight, (__half*)g, g_cs,
                                                                                   dweight, dbias, pixels, c);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

namespace {
__global__ void __launch_bounds__(256) grad_check_kernel(const float* g, long long n, float* scale_state) {
    bool bad = false;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) bad |= !isfinite(g[i]);
    if (__syncthreads_or(bad) && threadIdx.x == 0) scale_state[2] = 1.f;
}
__global__ void scale_update_kernel(float* st, float growth_interval, float max_scale, float min_scale) {
    if (st[2] != 0.f) { st[0] = fmaxf(st[0] * 0.5f, min_scale); st[1] = 0.f; st[2] = 0.f; return; }
    st[3] += 1.f;
    st[1] += 1.f;
    if (st[1] >= growth_interval) { st[0] = fminf(st[0] * 2.f, max_scale); st[1] = 0.f; }
}
}  // namespace

extern "C" int nrrt_grad_check(const float* grad, int64_t count, float* scale_state, void* stream) {
    NRRT_REQUIRE(grad && scale_state && count >= 1, "bad arguments");
    const long long blocks = (count + 255) / 256;
    grad_check_kernel<<<(int)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, (cudaStream_t)stream>>>(grad, count, scale_state);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_loss_scale_update(float* scale_state, int growth_interval, float max_scale, float min_scale, void* stream) {
    NRRT_REQUIRE(scale_state && growth_interval >= 1 && max_scale >= min_scale && min_scale > 0, "bad arguments");
    scale_update_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(scale_state, (float)growth_interval, max_scale, min_scale);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t count, int step, float lr,
                              float beta1, float beta2, float eps, float grad_scale, const float* scale_state, void* stream) {
    NRRT_REQUIRE(param && grad && exp_avg && exp_avg_sq && count >= 1 && step >= 1, "bad arguments");
    const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    const long long blocks = (count + 255) / 256;
    adam_kernel<<<(int)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, count, lr, beta1,
                                                                                              beta2, eps, (float)bc1, (float)sqrt(bc2), grad_scale, scale_state);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_to_half(const float* src, void* dst, int64_t count, void* stream) {
    NRRT_REQUIRE(src && dst && count >= 1, "bad arguments");
    const long long blocks = (count + 255) / 256;
    to_half_kernel<<<(int)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, (cudaStream_t)stream>>>(src, (__half*)dst, count);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_pack_dgrad_weight(const float* w, int cout, int taps, int ci_stride, int ci_rows, int co_cols, void* dst, void* stream) {
    NRRT_REQUIRE(w && dst && cout >= 1 && taps >= 1 && ci_stride >= 1 && ci_rows >= 1 && co_cols >= cout, "bad arguments");
    const long long total = (long long)ci_rows * taps * co_cols, blocks = (total + 255) / 256;
    pack_dgrad_kernel<<<(int)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, (cudaStream_t)stream>>>(w, (__half*)dst, cout, taps, ci_stride, ci_rows,
                                                                                                    co_cols);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_pack_dgrad_weights(const float* params, void* dst_base, const int64_t* descs_dev, int n_layers, void* stream) {
    NRRT_REQUIRE(params && dst_base && descs_dev && n_layers >= 1 && n_layers <= 65535, "bad arguments");
    pack_dgrad_batched_kernel<<<dim3(148, n_layers), 256, 0, (cudaStream_t)stream>>>(params, (__half*)dst_base, (const long long*)descs_dev);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_stem_wgrad(const float* x, const void* g, int64_t g_cs, int n, int cin, int h, int w, int cout, float* dweight,
                               float* dbias, void* stream) {
    NRRT_REQUIRE(x && g && dweight && n >= 1 && cin >= 1 && cin <= 4 && cout >= 1 && cout <= 64 && h >= 1 && w >= 1, "bad arguments");
    const int tiles = ((h + 15) / 16) * ((w + 15) / 16) * n;
    const size_t smem = ((size_t)cin * 18 * 18 + 256 * (size_t)cout) * sizeof(float);
    if (smem > 48 * 1024) NRRT_CUDA_TRY(cudaFuncSetAttribute(stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stem_wgrad_kernel<<<tiles, 256, smem, (cudaStream_t)stream>>>(x, (const __half*)g, g_cs, n, cin, h, w, cout, dweight, dbias);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_coords_fill_backward(const void* dcat, int64_t cs, int B, int h, int w, int c, float* dfeat, void* stream) {
    EW_CHECK(c);
    NRRT_REQUIRE(dcat && dfeat && B >= 1 && h >= 1 && w >= 1 && cs % 8 == 0, "bad arguments");
    int gx = ew_grid((long long)h * w, c);
    const int cap = (148 * 4 + B - 1) / B;       // about four blocks per SM over the batch; every block adds 4 c floats with atomics
    if (gx > cap) gx = cap;
    coords_fill_bwd_kernel<<<dim3(gx, B), kEw, ew_smem(c), (cudaStream_t)stream>>>((const __half*)dcat, cs, h, w, c, dfeat);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_coords_layer_backward(const float* dfeat_out, const float* f_out, const float* f_in, const float* weight, int B, int cin,
                                          int cout, float* dweight, float* dbias, float* dfeat_in, void* stream) {
    NRRT_REQUIRE(dfeat_out && f_out && f_in && weight && dweight && dbias && B >= 1 && cin >= 1 && cout >= 1, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const long long pairs = (long long)cin * cout;
    coords_wgrad_kernel<<<(int)((pairs + 255) / 256), 256, 0, st>>>(dfeat_out, f_out, f_in, B, cin, cout, dweight, dbias);
    if (dfeat_in) coords_dgrad_kernel<<<B * 4 * cin, 128, 0, st>>>(dfeat_out, f_out, weight, B, cin, cout, dfeat_in);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}
