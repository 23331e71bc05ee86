// sampler.cu — occupancy bit-packing, the neural sampler's CDF (built once per task) and the per-iteration sample draws.
//
// Replaces  generate_random_sample (rrt_star.py:63-68, ThreeD_rrt_star.py:64-70), generate_neural_sample
// (neural_rrt.py:73-99, ThreeD_neural_rrt.py:80-110), the sampler switch (neural_rrt.py:225-228) and the CNN-output
// post-processing (neural_rrt.py:350-355 + :64; ThreeD_neural_rrt.py:339-346 + :68).  Numeric contract: SURVEY.md A.8 —
// fp32 weights, NumPy's pairwise np.sum tree, strictly sequential fp32 np.cumsum, searchsorted compared in fp64.
//
// The reference rebuilds the CDF on every draw (O(N) per sample); the map never changes during a plan, so the CDF is a
// per-task constant: built once here, then each draw is a 32-ary warp search (4 dependent probes for 2^18 cells).
#include <algorithm>
#include <vector>

#include "nrrt_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------------------------------
// occupancy bit-packing: bit = 1 iff the reference treats the cell as free
// ------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void pack_kernel(const T* __restrict__ occ, int free_is_nonzero, int64_t cells, int words_per_map,
                            uint32_t* __restrict__ bits) {
    const int map = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int warps_per_grid = (gridDim.x * blockDim.x) >> 5;
    const T* src = occ + (size_t)map * cells;
    uint32_t* dst = bits + (size_t)map * words_per_map;
    for (int w0 = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; w0 < words_per_map; w0 += warps_per_grid * 32) {
        uint32_t mine = 0;
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
            const int64_t c = ((int64_t)(w0 + r) << 5) + lane;
            bool fr = false;
            if (c < cells) {
                const bool nz = src[c] != (T)0;
                fr = free_is_nonzero ? nz : !nz;
            }
            const uint32_t m = __ballot_sync(0xffffffffu, fr);
            if (r == lane) mine = m;
        }
        if (w0 + lane < words_per_map) dst[w0 + lane] = mine;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// CDF build.  stats[b] = {vmin, vmax, m (= min of h), S (= np.sum of w)}
// ------------------------------------------------------------------------------------------------------------------
struct LeafOp {
    int32_t a, b;  // leaf: (offset, len)   combine: (dst slot, src slot)
};

// planner-side heat value h_j from the raw network output (SURVEY a21 + a3/a4); all fp32, each op rounded separately
__device__ __forceinline__ float heat_value(float x, int mode, float vmin, float vmax) {
    if (mode == 0) return __fadd_rn(fminf(fmaxf(x, -3.0f), 1.0f), 10.0f);  // clamp(-3,1) ; heat_map[0] + 10
    if (mode == 1) {
        const float v = __fdiv_rn(__fsub_rn(x, vmin), __fsub_rn(vmax, vmin));  // (v - min) / (max - min) * 1.0
        const float masked = (v > 0.96f) ? v : 0.0f;                           // v[~(v > 0.96)] = 0
        return __fsub_rn(masked, 250.0f);                                      // heat_map - 250
    }
    return x;
}

__device__ __forceinline__ float block_reduce_min(float v, float* red, bool is_max) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float t = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, t) : fminf(v, t);
    }
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
    for (int w = 1; w < nw; ++w) r = is_max ? fmaxf(r, red[w]) : fminf(r, red[w]);
    return r;
}

// ---- exact parallel emulation of the sequential fp32 cumsum ---------------------------------------------------------
// np.cumsum is a strictly serial chain c_j = fl32(c_{j-1} + p_j) and fp32 addition is not associative, so a plain
// parallel scan gives different roundings.  But while c stays inside one binade [2^e, 2^(e+1)) it is an integer multiple
// K * u of u = ulp(c), K in [2^23, 2^24), and for p >= 0 round-to-nearest-even of c + p is integer arithmetic:
//     q = p / u (exact),  a = floor(q),  f = q - a:      K' = K + a + (f > 1/2)          if f != 1/2
//                                                        K' = K + a + ((K + a) & 1)       if f == 1/2 (tie -> even)
// so every element is a function K -> K + d[K & 1] described by the pair (d[0], d[1]) (equal unless the element is a
// tie), and these functions compose associatively: (f then g)[b] = f[b] + g[(b + f[b]) & 1].  A block-wide scan of the
// pairs therefore reproduces the serial chain bit for bit as long as K stays below 2^24.  The first element that leaves
// the binade is found with a block-wide min, evaluated with one real __fadd_rn, and the scan restarts behind it in the new
// binade (about 25 restarts per 2^18-cell map: c doubles that many times on its way from p_0 to ~1).  c = 0 and
// subnormal c are the same case with u = 2^-149.  Windows holding a NaN/Inf/negative weight (degenerate heat maps:
// S = 0 makes every p NaN) fall back to a serial loop on one thread.
constexpr int kFT = 1024;        // threads per CTA (one task per CTA at a time)
constexpr int kFV = 8;           // consecutive cells per thread per window
constexpr int kFW = kFT * kFV;   // cells per window
constexpr uint32_t kSat = 1u << 26;  // anything >= 2^24 means "left the binade"; saturate so sums never wrap

struct Pair2 {
    uint32_t d0, d1;
};
// correctly rounded w / S (see the Markstein note in cdf_fused_kernel); operands outside the safe range take __fdiv_rn
__device__ __forceinline__ float div_by(float w, float S, float rS, bool fastdiv) {
    if (w == 0.f && fastdiv) return 0.f;
    if (fastdiv && w >= 0x1p-60f && w <= 0x1p60f) {
        const float q = __fmul_rn(w, rS);
        const float r = __fmaf_rn(-q, S, w);
        return __fmaf_rn(r, rS, q);
    }
    return __fdiv_rn(w, S);
}
__device__ __forceinline__ Pair2 pair_then(const Pair2 f, const Pair2 g) {  // apply f, then g
    Pair2 r;
    r.d0 = min(f.d0 + ((f.d0 & 1u) ? g.d1 : g.d0), kSat);
    r.d1 = min(f.d1 + (((f.d1 + 1u) & 1u) ? g.d1 : g.d0), kSat);
    return r;
}
// the pair of one weight p >= 0 (finite): q = p * 2^(150 - ec) for a running sum with biased exponent field ec (>= 1; 1
// also covers 0 and subnormals).  The scale is applied as two exact power-of-two factors (2^149 is not a float); the
// float -> integer conversion saturates, so huge q simply read as "leaves the binade".
__device__ __forceinline__ void delta_of(float p, float scale_a, float scale_b, uint32_t& delta, bool& tie) {
    const float q = __fmul_rn(__fmul_rn(p, scale_a), scale_b);  // exact (power-of-two scaling) unless it over/underflows
    const uint32_t ai = min(__float2uint_rd(q), kSat);          // floor(q), saturated
    const float f = __fsub_rn(q, (float)ai);                    // exact fraction while q < 2^24 (beyond that: don't care)
    delta = ai + (f > 0.5f ? 1u : 0u);                          // round half down here; ties are patched by the caller
    tie = f == 0.5f;
}
__device__ __forceinline__ Pair2 pair_of(float p, float scale_a, float scale_b) {
    uint32_t d;
    bool tie;
    delta_of(p, scale_a, scale_b, d, tie);
    Pair2 r;
    r.d0 = d + ((tie && (d & 1u)) ? 1u : 0u);   // K even: K + a is odd iff a is odd -> round up to even
    r.d1 = d + ((tie && !(d & 1u)) ? 1u : 0u);  // K odd
    return r;
}

// One CTA per task (persistent over tasks, grid = #SMs so the rows being worked on stay L2-resident between phases):
//   A. min/max of the raw logits (3D) and min of the heat values                           [row read from HBM]
//   B. NumPy's pairwise sum of w = h - m with the exact tree: leaves of <= 128 elements with 8 strided accumulators
//      (8 lanes per leaf, the 8-way combine ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) is three xor-shuffles), then the binary
//      combine tree laid out by the host in `ops`                                           [row read from L2]
//   C. the exact parallel cumsum above over windows of 8192 cells, CDF written with streaming stores  [row read from L2]
__global__ void __launch_bounds__(kFT, 1) cdf_fused_kernel(const float* __restrict__ logits, int mode, int B, int64_t cells,
                                                           const LeafOp* __restrict__ leaves, int n_leaves,
                                                           const LeafOp* __restrict__ ops, const int32_t* __restrict__ level_off,
                                                           int n_levels, float4* __restrict__ stats, float* __restrict__ cdf) {
    extern __shared__ float leaf_sum[];  // [n_leaves]
    __shared__ float red[32];
    __shared__ Pair2 wtot[32];
    __shared__ float s_c;
    __shared__ int s_bad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const float* x = logits + (size_t)b * cells;
        float* out = cdf + (size_t)b * cells;
        float vmin_s = 0.f, vmax_s = 0.f, m_s = 0.f, S_s = 1.f;  // the task's statistics, as the serial fallback sees them
        const bool vec = (cells % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
        // the reference's arithmetic, serially on one thread (NaN/Inf propagate exactly as in NumPy)
        auto serial = [&](int64_t from, int64_t to, float c) {
            for (int64_t j = from; j < to; ++j) {
                const float pj = __fdiv_rn(__fsub_rn(heat_value(__ldg(x + j), mode, vmin_s, vmax_s), m_s), S_s);
                c = j == 0 ? pj : __fadd_rn(c, pj);
                out[j] = c;
            }
            return c;
        };

        // ---- A ----
        float vmin = 0.f, vmax = 0.f;
        if (mode == 1) {
            float lo = INFINITY, hi = -INFINITY;
            if (vec) {
                for (int64_t j = tid; j < cells / 4; j += blockDim.x) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + j);
                    lo = fminf(fminf(lo, fminf(v.x, v.y)), fminf(v.z, v.w));
                    hi = fmaxf(fmaxf(hi, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
                }
            } else {
                for (int64_t j = tid; j < cells; j += blockDim.x) { const float v = __ldg(x + j); lo = fminf(lo, v); hi = fmaxf(hi, v); }
            }
            vmin = block_reduce_min(lo, red, false);
            vmax = block_reduce_min(hi, red, true);
        }
        float m = INFINITY;
        if (vec) {
            for (int64_t j = tid; j < cells / 4; j += blockDim.x) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(x) + j);
                m = fminf(fminf(m, fminf(heat_value(v.x, mode, vmin, vmax), heat_value(v.y, mode, vmin, vmax))),
                          fminf(heat_value(v.z, mode, vmin, vmax), heat_value(v.w, mode, vmin, vmax)));
            }
        } else {
            for (int64_t j = tid; j < cells; j += blockDim.x) m = fminf(m, heat_value(__ldg(x + j), mode, vmin, vmax));
        }
        m = block_reduce_min(m, red, false);

        // ---- B ----
        const int k = tid & 7;
        for (int l0 = 0; l0 < n_leaves; l0 += blockDim.x / 8) {
            const int l = l0 + (tid >> 3);
            const bool on = l < n_leaves;
            const int off = on ? leaves[l].a : 0, len = on ? leaves[l].b : 0;
            const float* a = x + off;
            // every lane reaches the three shuffles at the same program point (leaves of one warp may differ in length)
            float r = 0.f;
            int i = 0;
            if (len >= 8) {
                r = __fsub_rn(heat_value(__ldg(a + k), mode, vmin, vmax), m);
                for (i = 8; i < len - (len % 8); i += 8) r = __fadd_rn(r, __fsub_rn(heat_value(__ldg(a + i + k), mode, vmin, vmax), m));
            }
            r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
            r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
            r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
            float res = r;
            if (len >= 8) {
                for (; i < len; ++i) res = __fadd_rn(res, __fsub_rn(heat_value(__ldg(a + i), mode, vmin, vmax), m));
            } else if (len > 0) {  // n < 8: plain left-to-right sum
                res = __fsub_rn(heat_value(__ldg(a), mode, vmin, vmax), m);
                for (int q = 1; q < len; ++q) res = __fadd_rn(res, __fsub_rn(heat_value(__ldg(a + q), mode, vmin, vmax), m));
            }
            if (on && k == 0) leaf_sum[l] = res;
        }
        __syncthreads();
        for (int lv = 0; lv < n_levels; ++lv) {  // deepest level first; left child's slot accumulates
            const int lo = level_off[lv], hi = level_off[lv + 1];
            for (int o = lo + tid; o < hi; o += blockDim.x) leaf_sum[ops[o].a] = __fadd_rn(leaf_sum[ops[o].a], leaf_sum[ops[o].b]);
            __syncthreads();
        }
        const float S = leaf_sum[0];
        vmin_s = vmin; vmax_s = vmax; m_s = m; S_s = S;
        // Markstein: with rS = RN(1/S), q = RN(w*rS), r = w - q*S (exact FMA), RN(q + r*rS) is the correctly rounded w / S
        // unless S's significand is all ones or something leaves the normal range; those cases use __fdiv_rn
        const float rS = __frcp_rn(S);
        const bool fastdiv = S >= 0x1p-60f && S <= 0x1p60f && (__float_as_uint(S) & 0x7fffffu) != 0x7fffffu;
        if (tid == 0) {
            if (stats) stats[b] = make_float4(vmin, vmax, m, S);
            s_c = 0.f;
        }
        __syncthreads();

        // ---- C ----
        for (int64_t w0 = 0; w0 < cells; w0 += kFW) {
            const int64_t wend = min(cells, w0 + (int64_t)kFW);
            const int64_t idx0 = w0 + (int64_t)tid * kFV;
            float p[kFV];
            if (vec && idx0 + kFV <= cells) {
#pragma unroll
                for (int h = 0; h < kFV / 4; ++h) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(x + idx0) + h);
                    p[4 * h + 0] = v.x; p[4 * h + 1] = v.y; p[4 * h + 2] = v.z; p[4 * h + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int v = 0; v < kFV; ++v) p[v] = idx0 + v < cells ? __ldg(x + idx0 + v) : 0.f;
            }
            bool odd = false;
#pragma unroll
            for (int v = 0; v < kFV; ++v) {
                p[v] = div_by(__fsub_rn(heat_value(p[v], mode, vmin, vmax), m), S, rS, fastdiv);
                if (idx0 + v < cells) odd |= !(p[v] >= 0.f && p[v] < INFINITY);
            }
            float c = s_c;
            if (__syncthreads_or(odd || !(c >= 0.f && c < INFINITY))) {
                if (tid == 0) s_c = serial(w0, wend, c);  // degenerate window
                __syncthreads();
                continue;
            }
            int64_t start = w0;
            for (;;) {
                const uint32_t cb = __float_as_uint(c);
                const uint32_t bec = cb >> 23;
                const uint32_t ec = bec ? bec : 1u;
                const uint32_t K0 = (cb & 0x7fffffu) | (bec ? 0x800000u : 0u);
                // 2^(150 - ec) = 2^ea * 2^eb, both factors normal floats
                const int sh = 150 - (int)ec;            // -104 .. 149
                const int ea = sh / 2, eb = sh - ea;
                const float scale_a = __uint_as_float((uint32_t)(ea + 127) << 23), scale_b = __uint_as_float((uint32_t)(eb + 127) << 23);
                if (tid == 0) s_bad = 0x7fffffff;
                // per-cell deltas; ties (fraction exactly 1/2, about 2^-17 of the cells) make the thread take the pair path
                uint32_t dl[kFV];
                bool any_tie = false;
                const bool live = idx0 + kFV > start && idx0 < wend;  // warp-uniform for all but one warp
                uint32_t sum = 0;
                if (live) {
#pragma unroll
                    for (int v = 0; v < kFV; ++v) {
                        const int64_t idx = idx0 + v;
                        bool tie;
                        delta_of(p[v], scale_a, scale_b, dl[v], tie);
                        if (!(idx >= start && idx < wend)) { dl[v] = 0; tie = false; }
                        any_tie |= tie;
                        sum = min(sum + dl[v], kSat);
                    }
                } else {
#pragma unroll
                    for (int v = 0; v < kFV; ++v) dl[v] = 0;
                }
                Pair2 agg = {sum, sum};
                if (any_tie) {
                    agg = Pair2{0u, 0u};
#pragma unroll
                    for (int v = 0; v < kFV; ++v) {
                        const int64_t idx = idx0 + v;
                        const Pair2 ev = (idx >= start && idx < wend) ? pair_of(p[v], scale_a, scale_b) : Pair2{0u, 0u};
                        agg = pair_then(agg, ev);
                    }
                }
                Pair2 inc = agg;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    Pair2 t;
                    t.d0 = __shfl_up_sync(0xffffffffu, inc.d0, o);
                    t.d1 = __shfl_up_sync(0xffffffffu, inc.d1, o);
                    if (lane >= o) inc = pair_then(t, inc);
                }
                Pair2 ex;
                ex.d0 = __shfl_up_sync(0xffffffffu, inc.d0, 1);
                ex.d1 = __shfl_up_sync(0xffffffffu, inc.d1, 1);
                if (lane == 0) ex = Pair2{0u, 0u};
                if (lane == 31) wtot[warp] = inc;
                __syncthreads();
                if (warp == 0) {
                    Pair2 t = wtot[lane];
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        Pair2 u;
                        u.d0 = __shfl_up_sync(0xffffffffu, t.d0, o);
                        u.d1 = __shfl_up_sync(0xffffffffu, t.d1, o);
                        if (lane >= o) t = pair_then(u, t);
                    }
                    Pair2 te;
                    te.d0 = __shfl_up_sync(0xffffffffu, t.d0, 1);
                    te.d1 = __shfl_up_sync(0xffffffffu, t.d1, 1);
                    if (lane == 0) te = Pair2{0u, 0u};
                    wtot[lane] = te;  // exclusive prefix of the warp totals
                }
                __syncthreads();
                const Pair2 pre = pair_then(wtot[warp], ex);
                uint32_t K = min(K0 + ((K0 & 1u) ? pre.d1 : pre.d0), kSat);
                uint32_t Kout[kFV];
                int bad = 0x7fffffff, bad_v = 0;
                uint32_t K_before_bad = 0;
                const uint32_t K_in = K;
                if (!any_tie) {
#pragma unroll
                    for (int v = 0; v < kFV; ++v) { K = min(K + dl[v], kSat); Kout[v] = K; }
                } else {
#pragma unroll
                    for (int v = 0; v < kFV; ++v) {
                        const int64_t idx = idx0 + v;
                        const Pair2 ev = (idx >= start && idx < wend) ? pair_of(p[v], scale_a, scale_b) : Pair2{0u, 0u};
                        K = min(K + ((K & 1u) ? ev.d1 : ev.d0), kSat);
                        Kout[v] = K;
                    }
                }
                if (live && K >= (1u << 24)) {  // K is monotone: only then can a cell of this thread have left the binade
                    uint32_t Kp = K_in;
#pragma unroll
                    for (int v = 0; v < kFV; ++v) {
                        const int64_t idx = idx0 + v;
                        if (idx >= start && idx < wend && Kout[v] >= (1u << 24) && bad == 0x7fffffff) {
                            bad = (int)(idx - w0); bad_v = v; K_before_bad = Kp;
                        }
                        Kp = Kout[v];
                    }
                }
                int wbad = bad;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) wbad = min(wbad, __shfl_xor_sync(0xffffffffu, wbad, o));
                if (lane == 0 && wbad != 0x7fffffff) atomicMin(&s_bad, wbad);
                __syncthreads();
                const int gbad = s_bad;  // window-relative index of the first element that leaves the binade
                const int64_t stop = gbad == 0x7fffffff ? wend : w0 + gbad;
                // finished elements [start, stop)
                auto value_of = [&](uint32_t Kv) { return __uint_as_float(Kv >= 0x800000u ? ((ec << 23) | (Kv & 0x7fffffu)) : Kv); };
                if (vec && idx0 >= start && idx0 + kFV <= stop) {
#pragma unroll
                    for (int h = 0; h < kFV / 4; ++h)
                        __stcs(reinterpret_cast<float4*>(out + idx0) + h,
                               make_float4(value_of(Kout[4 * h]), value_of(Kout[4 * h + 1]), value_of(Kout[4 * h + 2]), value_of(Kout[4 * h + 3])));
                } else {
#pragma unroll
                    for (int v = 0; v < kFV; ++v)
                        if (idx0 + v >= start && idx0 + v < stop) out[idx0 + v] = value_of(Kout[v]);
                }
                if (gbad == 0x7fffffff) {
                    // the thread that owns the window's last cell publishes the running sum
                    if (wend - 1 >= idx0 && wend - 1 < idx0 + kFV && wend - 1 >= start) s_c = value_of(Kout[(int)(wend - 1 - idx0)]);
                    __syncthreads();
                    break;
                }
                if (bad == gbad) {  // owner: one real fp32 addition carries the sum into its new binade
                    float pb = 0.f;
#pragma unroll
                    for (int v = 0; v < kFV; ++v) if (v == bad_v) pb = p[v];
                    const float cn = __fadd_rn(value_of(K_before_bad), pb);
                    out[w0 + gbad] = cn;
                    s_c = cn;
                }
                __syncthreads();
                c = s_c;
                start = w0 + gbad + 1;
                if (start >= wend) break;
                if (!(c < INFINITY)) {  // the sum overflowed: finish the window serially, later windows take the serial path
                    if (tid == 0) s_c = serial(start, wend, c);
                    break;
                }
            }
            __syncthreads();
        }
        __syncthreads();
    }
}

// ---- cdf_stream_kernel: the same three phases as cdf_fused_kernel, fed by a bulk-copy (TMA) pipeline ------------------
// The fused kernel above exposes a global-load latency per window.  Here warp 0 is a producer: one thread streams the
// task's row through a ring of kStages x 32 KB shared-memory stages with cp.async.bulk (1-D TMA) + mbarrier complete_tx,
// once per phase ([min/max,] min, pairwise sum, cumsum).  The order of the copies does not depend on any result, so the
// producer runs ahead across windows, phases and tasks and the 16 consumer warps never wait on HBM/L2 latency.  A chunk
// is a run of whole pairwise-sum leaves (<= 8192 cells), so every phase uses the same chunking and leaves never
// straddle a stage.  Needs cells % 4 == 0 and 16-byte aligned rows (bulk-copy alignment); other shapes use cdf_fused_kernel.
constexpr int kSC = 512;                 // consumer threads
constexpr int kSV = 16;                  // cells per consumer thread per chunk
constexpr int kChunk = kSC * kSV;        // 8192 cells = 32 KB per stage
constexpr int kStages = 5;

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void sbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void sbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "SW_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra SW_DONE;\n\t"
        "bra SW_LOOP;\n\t"
        "SW_DONE:\n\t}" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kSC) : "memory"); }

// min (or max) over the 512 consumer threads; every consumer thread gets the result
__device__ __forceinline__ float consumer_reduce(float v, float* red, bool is_max, int ct) {
    const int lane = ct & 31, warp = ct >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float t = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, t) : fminf(v, t);
    }
    consumer_sync();
    if (lane == 0) red[warp] = v;
    consumer_sync();
    float r = red[0];
#pragma unroll
    for (int w = 1; w < kSC / 32; ++w) r = is_max ? fmaxf(r, red[w]) : fminf(r, red[w]);
    return r;
}

__global__ void __launch_bounds__(kSC + 32, 1) cdf_stream_kernel(const float* __restrict__ logits, int mode, int B, int64_t cells,
                                                                 const LeafOp* __restrict__ leaves, int n_leaves,
                                                                 const LeafOp* __restrict__ ops,
                                                                 const int32_t* __restrict__ level_off, int n_levels,
                                                                 const int32_t* __restrict__ chunk_leaf, int n_chunks,
                                                                 float4* __restrict__ stats, float* __restrict__ cdf) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float* ring = reinterpret_cast<float*>(smem_raw);                      // [kStages][kChunk]
    float* leaf_sum = ring + kStages * kChunk;                              // [n_leaves]
    __shared__ uint64_t bars[2 * kStages];
    __shared__ float red[kSC / 32];
    __shared__ Pair2 wtot[kSC / 32];
    __shared__ float s_c;
    __shared__ int s_bad;
    const uint32_t bar0 = s_u32(bars);
    auto full = [&](int st) { return bar0 + 8u * st; };
    auto empty = [&](int st) { return bar0 + 8u * (kStages + st); };
    const int n_phases = mode == 1 ? 4 : 3;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { sbar_init(full(i), 1); sbar_init(empty(i), kSC / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (threadIdx.x < 32) {
        // ===== producer =====
        if (threadIdx.x == 0) {
            int st = 0;
            uint32_t ph = 0;
            for (int b = blockIdx.x; b < B; b += gridDim.x) {
                const float* x = logits + (size_t)b * cells;
                for (int phase = 0; phase < n_phases; ++phase)
                    for (int c = 0; c < n_chunks; ++c) {
                        const int off = __ldg(&leaves[__ldg(chunk_leaf + c)].a);
                        const int end = c + 1 < n_chunks ? __ldg(&leaves[__ldg(chunk_leaf + c + 1)].a) : (int)cells;
                        const uint32_t bytes = (uint32_t)(end - off) * 4u;
                        sbar_wait(empty(st), ph ^ 1u);
                        sbar_expect_tx(full(st), bytes);
                        bulk_load(s_u32(ring + st * kChunk), x + off, bytes, full(st));
                        if (++st == kStages) { st = 0; ph ^= 1u; }
                    }
            }
        }
        return;
    }

    // ===== consumers =====
    const int ct = threadIdx.x - 32, lane = ct & 31, warp = ct >> 5;
    int st = 0;
    uint32_t ph = 0;
    auto release = [&]() {
        __syncwarp();
        if (lane == 0) sbar_arrive(empty(st));
        if (++st == kStages) { st = 0; ph ^= 1u; }
    };
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const float* x = logits + (size_t)b * cells;
        float* out = cdf + (size_t)b * cells;
        float vmin = 0.f, vmax = 0.f;
        // ---- min / max of the raw logits (3D normalisation) ----
        if (mode == 1) {
            float lo = INFINITY, hi = -INFINITY;
            for (int c = 0; c < n_chunks; ++c) {
                const int off = __ldg(&leaves[__ldg(chunk_leaf + c)].a);
                const int end = c + 1 < n_chunks ? __ldg(&leaves[__ldg(chunk_leaf + c + 1)].a) : (int)cells;
                sbar_wait(full(st), ph);
                const float4* src = reinterpret_cast<const float4*>(ring + st * kChunk);
                for (int j = ct; j < (end - off) / 4; j += kSC) {
                    const float4 v = src[j];
                    lo = fminf(fminf(lo, fminf(v.x, v.y)), fminf(v.z, v.w));
                    hi = fmaxf(fmaxf(hi, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
                }
                release();
            }
            vmin = consumer_reduce(lo, red, false, ct);
            vmax = consumer_reduce(hi, red, true, ct);
        }
        // ---- A: min of the heat values ----
        float m = INFINITY;
        for (int c = 0; c < n_chunks; ++c) {
            const int off = __ldg(&leaves[__ldg(chunk_leaf + c)].a);
            const int end = c + 1 < n_chunks ? __ldg(&leaves[__ldg(chunk_leaf + c + 1)].a) : (int)cells;
            sbar_wait(full(st), ph);
            const float4* src = reinterpret_cast<const float4*>(ring + st * kChunk);
            for (int j = ct; j < (end - off) / 4; j += kSC) {
                const float4 v = src[j];
                m = fminf(fminf(m, fminf(heat_value(v.x, mode, vmin, vmax), heat_value(v.y, mode, vmin, vmax))),
                          fminf(heat_value(v.z, mode, vmin, vmax), heat_value(v.w, mode, vmin, vmax)));
            }
            release();
        }
        m = consumer_reduce(m, red, false, ct);
        // ---- B: NumPy's pairwise sum (same leaf code as cdf_fused_kernel, operands in shared memory) ----
        const int k = ct & 7;
        for (int c = 0; c < n_chunks; ++c) {
            const int l_begin = __ldg(chunk_leaf + c), l_end = c + 1 < n_chunks ? __ldg(chunk_leaf + c + 1) : n_leaves;
            const int off0 = __ldg(&leaves[l_begin].a);
            sbar_wait(full(st), ph);
            const float* src = ring + st * kChunk;
            for (int l0 = l_begin; l0 < l_end; l0 += kSC / 8) {
                const int l = l0 + (ct >> 3);
                const bool on = l < l_end;
                const int off = on ? __ldg(&leaves[l].a) - off0 : 0, len = on ? __ldg(&leaves[l].b) : 0;
                const float* a = src + off;
                float r = 0.f;
                int i = 0;
                if (len >= 8) {
                    r = __fsub_rn(heat_value(a[k], mode, vmin, vmax), m);
                    for (i = 8; i < len - (len % 8); i += 8) r = __fadd_rn(r, __fsub_rn(heat_value(a[i + k], mode, vmin, vmax), m));
                }
                r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
                r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
                r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
                float res = r;
                if (len >= 8) {
                    for (; i < len; ++i) res = __fadd_rn(res, __fsub_rn(heat_value(a[i], mode, vmin, vmax), m));
                } else if (len > 0) {
                    res = __fsub_rn(heat_value(a[0], mode, vmin, vmax), m);
                    for (int q = 1; q < len; ++q) res = __fadd_rn(res, __fsub_rn(heat_value(a[q], mode, vmin, vmax), m));
                }
                if (on && k == 0) leaf_sum[l] = res;
            }
            release();
        }
        consumer_sync();
        for (int lv = 0; lv < n_levels; ++lv) {
            const int lo = level_off[lv], hi = level_off[lv + 1];
            for (int o = lo + ct; o < hi; o += kSC) leaf_sum[ops[o].a] = __fadd_rn(leaf_sum[ops[o].a], leaf_sum[ops[o].b]);
            consumer_sync();
        }
        const float S = leaf_sum[0];
        const float rS = __frcp_rn(S);
        const bool fastdiv = S >= 0x1p-60f && S <= 0x1p60f && (__float_as_uint(S) & 0x7fffffu) != 0x7fffffu;
        const bool regular = S > 0.f && S < INFINITY && fabsf(m) < INFINITY;  // else: NumPy's NaN/Inf arithmetic, serially
        if (ct == 0) {
            if (stats) stats[b] = make_float4(vmin, vmax, m, S);
            s_c = 0.f;
        }
        consumer_sync();  // also: every thread has read leaf_sum[0] before the next task overwrites it

        // ---- C: exact parallel cumsum, one chunk = one window ----
        for (int c = 0; c < n_chunks; ++c) {
            const int off = __ldg(&leaves[__ldg(chunk_leaf + c)].a);
            const int end = c + 1 < n_chunks ? __ldg(&leaves[__ldg(chunk_leaf + c + 1)].a) : (int)cells;
            const int wlen = end - off;
            sbar_wait(full(st), ph);
            const float* src = ring + st * kChunk;
            const int i0 = ct * kSV;  // window-relative index of this thread's first cell
            float p[kSV];
#pragma unroll
            for (int h = 0; h < kSV / 4; ++h) {
                const float4 v = i0 + 4 * h < wlen ? *reinterpret_cast<const float4*>(src + i0 + 4 * h) : make_float4(0.f, 0.f, 0.f, 0.f);
                p[4 * h + 0] = v.x; p[4 * h + 1] = v.y; p[4 * h + 2] = v.z; p[4 * h + 3] = v.w;
            }
            float cprev = s_c;
            if (!regular || !(cprev >= 0.f && cprev < INFINITY)) {
                if (ct == 0) {
                    float cc = cprev;
                    for (int j = off; j < end; ++j) {
                        const float pj = __fdiv_rn(__fsub_rn(heat_value(__ldg(x + j), mode, vmin, vmax), m), S);
                        cc = j == 0 ? pj : __fadd_rn(cc, pj);
                        out[j] = cc;
                    }
                    s_c = cc;
                }
                release();
                consumer_sync();
                continue;
            }
#pragma unroll
            for (int v = 0; v < kSV; ++v) p[v] = div_by(__fsub_rn(heat_value(p[v], mode, vmin, vmax), m), S, rS, fastdiv);
            int start = 0;
            float cur = cprev;
            for (;;) {
                const uint32_t cb = __float_as_uint(cur);
                const uint32_t bec = cb >> 23;
                const uint32_t ec = bec ? bec : 1u;
                const uint32_t K0 = (cb & 0x7fffffu) | (bec ? 0x800000u : 0u);
                const int sh = 150 - (int)ec;
                const int ea = sh / 2, eb = sh - ea;
                const float scale_a = __uint_as_float((uint32_t)(ea + 127) << 23), scale_b = __uint_as_float((uint32_t)(eb + 127) << 23);
                if (ct == 0) s_bad = 0x7fffffff;
                uint32_t dl[kSV];
                bool any_tie = false;
                const bool live = i0 + kSV > start && i0 < wlen;
                uint32_t sum = 0;
                if (live) {
#pragma unroll
                    for (int v = 0; v < kSV; ++v) {
                        bool tie;
                        delta_of(p[v], scale_a, scale_b, dl[v], tie);
                        if (!(i0 + v >= start && i0 + v < wlen)) { dl[v] = 0; tie = false; }
                        any_tie |= tie;
                        sum = min(sum + dl[v], kSat);
                    }
                } else {
#pragma unroll
                    for (int v = 0; v < kSV; ++v) dl[v] = 0;
                }
                Pair2 agg = {sum, sum};
                if (any_tie) {
                    agg = Pair2{0u, 0u};
#pragma unroll
                    for (int v = 0; v < kSV; ++v) {
                        const Pair2 ev = (i0 + v >= start && i0 + v < wlen) ? pair_of(p[v], scale_a, scale_b) : Pair2{0u, 0u};
                        agg = pair_then(agg, ev);
                    }
                }
                Pair2 inc = agg;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    Pair2 t;
                    t.d0 = __shfl_up_sync(0xffffffffu, inc.d0, o);
                    t.d1 = __shfl_up_sync(0xffffffffu, inc.d1, o);
                    if (lane >= o) inc = pair_then(t, inc);
                }
                Pair2 ex;
                ex.d0 = __shfl_up_sync(0xffffffffu, inc.d0, 1);
                ex.d1 = __shfl_up_sync(0xffffffffu, inc.d1, 1);
                if (lane == 0) ex = Pair2{0u, 0u};
                if (lane == 31) wtot[warp] = inc;
                consumer_sync();
                if (warp == 0) {
                    Pair2 t = lane < kSC / 32 ? wtot[lane] : Pair2{0u, 0u};
#pragma unroll
                    for (int o = 1; o < kSC / 32; o <<= 1) {
                        Pair2 u;
                        u.d0 = __shfl_up_sync(0xffffffffu, t.d0, o);
                        u.d1 = __shfl_up_sync(0xffffffffu, t.d1, o);
                        if (lane >= o) t = pair_then(u, t);
                    }
                    Pair2 te;
                    te.d0 = __shfl_up_sync(0xffffffffu, t.d0, 1);
                    te.d1 = __shfl_up_sync(0xffffffffu, t.d1, 1);
                    if (lane == 0) te = Pair2{0u, 0u};
                    if (lane < kSC / 32) wtot[lane] = te;
                }
                consumer_sync();
                const Pair2 pre = pair_then(wtot[warp], ex);
                uint32_t K = min(K0 + ((K0 & 1u) ? pre.d1 : pre.d0), kSat);
                const uint32_t K_in = K;
                uint32_t Kout[kSV];
                if (!any_tie) {
#pragma unroll
                    for (int v = 0; v < kSV; ++v) { K = min(K + dl[v], kSat); Kout[v] = K; }
                } else {
#pragma unroll
                    for (int v = 0; v < kSV; ++v) {
                        const Pair2 ev = (i0 + v >= start && i0 + v < wlen) ? pair_of(p[v], scale_a, scale_b) : Pair2{0u, 0u};
                        K = min(K + ((K & 1u) ? ev.d1 : ev.d0), kSat);
                        Kout[v] = K;
                    }
                }
                int bad = 0x7fffffff, bad_v = 0;
                uint32_t K_before_bad = 0;
                if (live && K >= (1u << 24)) {
                    uint32_t Kp = K_in;
#pragma unroll
                    for (int v = 0; v < kSV; ++v) {
                        if (i0 + v >= start && i0 + v < wlen && Kout[v] >= (1u << 24) && bad == 0x7fffffff) {
                            bad = i0 + v; bad_v = v; K_before_bad = Kp;
                        }
                        Kp = Kout[v];
                    }
                }
                int wbad = bad;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) wbad = min(wbad, __shfl_xor_sync(0xffffffffu, wbad, o));
                if (lane == 0 && wbad != 0x7fffffff) atomicMin(&s_bad, wbad);
                consumer_sync();
                const int gbad = s_bad;
                const int stop = gbad == 0x7fffffff ? wlen : gbad;
                auto value_of = [&](uint32_t Kv) { return __uint_as_float(Kv >= 0x800000u ? ((ec << 23) | (Kv & 0x7fffffu)) : Kv); };
                float* o = out + off + i0;
                if (i0 >= start && i0 + kSV <= stop) {
#pragma unroll
                    for (int h = 0; h < kSV / 4; ++h)
                        __stcs(reinterpret_cast<float4*>(o) + h, make_float4(value_of(Kout[4 * h]), value_of(Kout[4 * h + 1]),
                                                                             value_of(Kout[4 * h + 2]), value_of(Kout[4 * h + 3])));
                } else {
#pragma unroll
                    for (int v = 0; v < kSV; ++v)
                        if (i0 + v >= start && i0 + v < stop) o[v] = value_of(Kout[v]);
                }
                if (gbad == 0x7fffffff) {
                    if (wlen - 1 >= i0 && wlen - 1 < i0 + kSV) {
                        float last = 0.f;
#pragma unroll
                        for (int v = 0; v < kSV; ++v) if (v == wlen - 1 - i0) last = value_of(Kout[v]);
                        s_c = last;
                    }
                    consumer_sync();
                    break;
                }
                if (bad == gbad) {
                    float pb = 0.f;
#pragma unroll
                    for (int v = 0; v < kSV; ++v) if (v == bad_v) pb = p[v];
                    const float cn = __fadd_rn(value_of(K_before_bad), pb);
                    out[off + gbad] = cn;
                    s_c = cn;
                }
                consumer_sync();
                cur = s_c;
                start = gbad + 1;
                if (start >= wlen) break;
                if (!(cur < INFINITY)) {  // the sum overflowed: finish the window serially
                    if (ct == 0) {
                        float cc = cur;
                        for (int j = off + start; j < end; ++j) {
                            cc = __fadd_rn(cc, __fdiv_rn(__fsub_rn(heat_value(__ldg(x + j), mode, vmin, vmax), m), S));
                            out[j] = cc;
                        }
                        s_c = cc;
                    }
                    break;
                }
            }
            // The stage is released only here, after the window: releasing right after the register loads let the
            // producer's refill of this stage show up in the loaded values on B200 (measured; ptxas may re-materialise
            // shared-memory loads across the arrive).  The producer is still up to kStages - 1 windows ahead.
            release();
            consumer_sync();
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// per-iteration sample draws, one warp per task
// ------------------------------------------------------------------------------------------------------------------
struct DrawArgs {
    NrrtPlanParams p;
    const uint32_t* occ_bits;
    const int32_t* task_to_map;
    const float* cdf;
    const int32_t* task_ids;
    int B;
    int16_t* samples;
    int8_t* neural_flag;
    int64_t* n_draws;
    int32_t* draw_status;
    int words_per_map;
    int64_t cells;
};

// inverse-CDF lookup = first index j in [0, n] with !(cdf[j] < u) (np.searchsorted side='left', fp64 compare, NaN sorts last)
template <int DIM>
__global__ void __launch_bounds__(128) draw_kernel(const DrawArgs a) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= a.B) return;
    const uint32_t* grid = a.occ_bits + (size_t)a.task_to_map[b] * a.words_per_map;
    const float* cdf = a.cdf ? a.cdf + (size_t)b * a.cells : nullptr;
    const uint32_t stream_id = (uint32_t)a.task_ids[b];
    const uint64_t seed = a.p.seed;
    const int s0 = a.p.shape[0], s1 = a.p.shape[1], s2 = DIM == 3 ? a.p.shape[2] : 1;
    const int max_iter = a.p.max_iterations;
    int16_t* out = a.samples + (size_t)b * max_iter * DIM;

    // Draws are generated 4 x 32 at a time (lane l holds draws 32 j + l of the group, j = 0..3) and consumed in order.
    // Speculative inverse-CDF lookup: which draws of the stream are neural samples is only known as the stream is consumed
    // (bias test, rejection retries), but a lookup has no side effect, so every lane searches ITS four uniforms when the
    // group is generated: 128 binary searches in flight (four independent chains per lane, 19 dependent probes per group)
    // instead of one 32-ary warp search (4 dependent probes) per neural draw as it comes up.  The kernel is a chain of
    // DRAM / L2 round trips (one warp per task); this cut the 3D draw pass from 121 ms to a quarter.
    // Same partition point as np.searchsorted(side='left'): first j with !(cdf[j] < u).
    constexpr int NB = 4;
    uint64_t k = a.p.draw_start, grp = ~0ull;
    double u_lane[NB] = {0.0, 0.0, 0.0, 0.0};
    int idx_lane[NB] = {0, 0, 0, 0};
    uint32_t free_lane = 0u;  // bit j: the cell of lookup j is free
    const uint32_t* grid_words = grid;
    auto gen_group = [&]() {  // generate (and look up) the group of 128 draws that holds draw k, if it is not the current one
        if ((k >> 7) != grp) {
            grp = k >> 7;
#pragma unroll
            for (int j = 0; j < NB; ++j) u_lane[j] = nrrt_uniform(seed, stream_id, (grp << 7) + (uint64_t)(j * 32 + lane));
            if (cdf) {
                int lo[NB], hi[NB];
#pragma unroll
                for (int j = 0; j < NB; ++j) { lo[j] = 0; hi[j] = (int)a.cells; }
                bool more = a.cells > 0;
                while (more) {
                    float v[NB];
                    int mid[NB];
#pragma unroll
                    for (int j = 0; j < NB; ++j) {  // the four loads are independent: issued together
                        mid[j] = (lo[j] + hi[j]) >> 1;
                        v[j] = lo[j] < hi[j] ? __ldg(cdf + mid[j]) : 0.f;
                    }
                    more = false;
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        if (lo[j] < hi[j]) {
                            if ((double)v[j] < u_lane[j]) lo[j] = mid[j] + 1; else hi[j] = mid[j];
                        }
                        more = more || lo[j] < hi[j];
                    }
                }
#pragma unroll
                for (int j = 0; j < NB; ++j) idx_lane[j] = lo[j];
                if (DIM == 3) {  // ... and the occupancy test of the cell each lookup landed on (3D redraws until free)
                    uint32_t wv[NB];
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        const int64_t c = lo[j] > (int)a.cells - 1 ? a.cells - 1 : (int64_t)lo[j];
                        wv[j] = __ldg(grid_words + (c >> 5)) >> (c & 31);
                    }
                    free_lane = 0u;
#pragma unroll
                    for (int j = 0; j < NB; ++j) free_lane |= (wv[j] & 1u) << j;
                }
            }
        }
    };
    auto next_u = [&]() -> double {
        gen_group();
        const int j = (int)((k >> 5) & 3);
        const double mine = j == 0 ? u_lane[0] : (j == 1 ? u_lane[1] : (j == 2 ? u_lane[2] : u_lane[3]));
        const double u = __shfl_sync(0xffffffffu, mine, (int)(k & 31));
        ++k;
        return u;
    };
    auto lookup_of = [&](uint64_t kk) -> int {  // speculative lookup of draw kk (of the current group)
        const int j = (int)((kk >> 5) & 3);
        const int mine = j == 0 ? idx_lane[0] : (j == 1 ? idx_lane[1] : (j == 2 ? idx_lane[2] : idx_lane[3]));
        return __shfl_sync(0xffffffffu, mine, (int)(kk & 31));
    };
    auto lookup_free = [&](uint64_t kk) -> bool {
        return (__shfl_sync(0xffffffffu, free_lane, (int)(kk & 31)) >> ((kk >> 5) & 3)) & 1u;
    };
    auto randint0 = [&](int n) -> int {  // random.randint(0, n - 1) on the shared stream
        const int v = __double2int_rd(__dmul_rn(next_u(), (double)n));
        return min(v, n - 1);
    };
    auto is_free = [&](int64_t idx) -> bool { return (__ldg(grid + (idx >> 5)) >> (idx & 31)) & 1u; };

    // After 256 redraws in a row: decide once whether ANY cell the lookup can return is free (the reference's `while True`
    // would spin forever otherwise, ThreeD_neural_rrt.py:82-110).  A cell is reachable when it carries probability mass
    // (cdf[i] > cdf[i-1]); the first and the last cell are treated as reachable (u = 0 / u above the rounded total land
    // there), so the test never gives up on a task that could succeed.
    auto any_reachable_free = [&]() -> bool {
        bool any = false;
        const int n = (int)a.cells;
#pragma unroll 4
        for (int i = lane; i < n; i += 32) {
            const float ci = __ldg(cdf + i);
            const float pv = i ? __ldg(cdf + i - 1) : -1.f;
            if ((ci > pv || i == n - 1) && is_free(i)) any = true;
        }
        return __any_sync(0xffffffffu, any);
    };

    int status = 0;
    bool checked = false;
    int16_t keep[DIM];
    int8_t keep_flag = 0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) keep[d] = 0;
    int it = 0;
    for (; it < max_iter && status == 0; ++it) {
        bool use_neural = false;
        if (a.p.sampler_mode == 0) {
            if (cdf) use_neural = next_u() < a.p.neural_bias;  // random.random() < self.neural_bias
        } else {
            use_neural = a.p.sampler_mode == 1;                // the sampler method called directly: no bias draw
        }
        int c[3] = {0, 0, 0};
        int rej = 0;
        for (;;) {
            if (use_neural) {
                if (DIM == 3 && rej > 0) {
                    // inside a run of redraws every draw is a neural redraw until one lands on a free cell, and the verdicts
                    // of the current group are already known: skip its failures at once (same draws consumed, same count)
                    gen_group();
                    const int p0 = (int)(k & 127);
                    int pos = -1;
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        uint32_t m = __ballot_sync(0xffffffffu, (free_lane >> j) & 1u);
                        if (j == (p0 >> 5)) m &= 0xffffffffu << (p0 & 31);
                        if (pos < 0 && j >= (p0 >> 5) && m) pos = j * 32 + __ffs(m) - 1;
                    }
                    const int fails = (pos >= 0 ? pos : 128) - p0;  // failed redraws before the next success / the group's end
                    const int room = a.p.max_reject - rej;          // failures still allowed
                    if (fails >= room) { k += (uint64_t)room; status = NRRT_SAMPLER_STUCK; break; }
                    k += (uint64_t)fails;
                    rej += fails;
                    if (rej >= 256 && !checked) {  // (the draw count reported for a stuck task is where the scan ran)
                        checked = true;
                        if (!any_reachable_free()) { status = NRRT_SAMPLER_STUCK; break; }
                    }
                    if (pos < 0) continue;  // nothing free in the rest of this group: on to the next one
                }
                const uint64_t kk = k;  // this draw's position in the stream
                (void)next_u();         // random.uniform(0, 1) = 0 + (1 - 0) * random()   (generates kk's group if needed)
                int64_t idx = lookup_of(kk);
                idx = idx > a.cells - 1 ? a.cells - 1 : idx;  // min(max(index, 0), N - 1)
                if (DIM == 2) { c[0] = (int)(idx / s1); c[1] = (int)(idx % s1); break; }
                const int64_t yz = (int64_t)s1 * s2;
                c[0] = (int)(idx / yz); c[1] = (int)((idx % yz) / s2); c[2] = (int)(idx % s2);
                if (lookup_free(kk)) break;  // == is_free(idx), tested when the group was generated
            } else if (DIM == 2) {
                const int x = randint0(s1);  // x = randint(0, map_width - 1)
                const int y = randint0(s0);  // y = randint(0, map_height - 1)
                if (is_free((int64_t)y * s1 + x)) { c[0] = y; c[1] = x; break; }
            } else {
                const int x = randint0(s1);  // map_width = shape[1] (sic), indexes axis 0
                const int y = randint0(s0);
                const int z = randint0(s2);
                const int xi = min(x, s0 - 1), yi = min(y, s1 - 1);
                if (is_free(((int64_t)xi * s1 + yi) * s2 + z)) { c[0] = x; c[1] = y; c[2] = z; break; }
            }
            if (++rej >= a.p.max_reject) { status = NRRT_SAMPLER_STUCK; break; }
            if (DIM == 3 && use_neural && rej >= 256 && !checked) {
                checked = true;
                if (!any_reachable_free()) { status = NRRT_SAMPLER_STUCK; break; }
            }
        }
        if (lane == (it & 31)) {
#pragma unroll
            for (int d = 0; d < DIM; ++d) keep[d] = (int16_t)c[d];
            keep_flag = (int8_t)use_neural;
        }
        if ((it & 31) == 31 || it == max_iter - 1 || status != 0) {  // flush up to 32 samples, one per lane
            const int base = it & ~31;
            if (base + lane <= it) {
#pragma unroll
                for (int d = 0; d < DIM; ++d) out[(size_t)(base + lane) * DIM + d] = keep[d];
                if (a.neural_flag) a.neural_flag[(size_t)b * max_iter + base + lane] = keep_flag;
            }
        }
    }
    if (lane == 0) {
        a.draw_status[b] = status;
        if (a.n_draws) a.n_draws[b] = (int64_t)k;
    }
}

// NumPy pairwise-sum tree for n elements (numpy/_core/src/umath/loops_utils.h.src pairwise_sum):
// leaves in array order; combine ops grouped by depth, deepest first; every op adds the right child's slot into the
// left child's slot, so slot 0 ends up holding the root.
struct Tree {
    std::vector<LeafOp> leaves;
    std::vector<std::vector<LeafOp>> by_depth;
    int build(int64_t off, int64_t n, int depth) {  // returns the slot holding this subtree's sum
        if (n <= 128) {
            leaves.push_back({(int32_t)off, (int32_t)n});
            return (int)leaves.size() - 1;
        }
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        const int l = build(off, n2, depth + 1);
        const int r = build(off + n2, n - n2, depth + 1);
        if ((int)by_depth.size() <= depth) by_depth.resize(depth + 1);
        by_depth[depth].push_back({l, r});
        return l;
    }
};

}  // namespace

extern "C" size_t nrrt_words_per_map(int64_t cells) { return (size_t)((((cells + 31) >> 5) + 3) & ~(int64_t)3); }

extern "C" int nrrt_pack_occupancy(const void* occ, int dtype, int free_is_nonzero, int n_maps, int64_t cells,
                                   uint32_t* bits, void* stream) {
    NRRT_REQUIRE(occ && bits, "null pointer");
    NRRT_REQUIRE(n_maps >= 0 && cells > 0 && cells < ((int64_t)1 << 31), "bad sizes");
    NRRT_REQUIRE(dtype >= 0 && dtype <= 2, "dtype must be 0 (u8), 1 (f32) or 2 (f64)");
    if (n_maps == 0) return NRRT_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int wpm = (int)nrrt_words_per_map(cells);
    const int warps_needed = (wpm + 31) / 32;
    dim3 grid((unsigned)std::min(64, (warps_needed + 7) / 8), (unsigned)n_maps);
    if (dtype == 0) pack_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t*)occ, free_is_nonzero, cells, wpm, bits);
    else if (dtype == 1) pack_kernel<float><<<grid, 256, 0, st>>>((const float*)occ, free_is_nonzero, cells, wpm, bits);
    else pack_kernel<double><<<grid, 256, 0, st>>>((const double*)occ, free_is_nonzero, cells, wpm, bits);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" size_t nrrt_cdf_workspace_bytes(int B, int64_t cells) {
    const size_t max_leaves = (size_t)(cells / 64 + 2);  // every leaf of a split has > 64 elements
    return align_up((size_t)B * sizeof(float4), 256) + align_up(max_leaves * sizeof(LeafOp), 256) * 2 + 256 * sizeof(int32_t) +
           align_up(max_leaves * sizeof(int32_t), 256);  // stats | leaves | ops | level offsets | chunk table
}

extern "C" int nrrt_cdf_build(const float* logits, int mode, int B, int64_t cells, float* cdf, void* workspace,
                              size_t ws_bytes, void* stream) {
    NRRT_REQUIRE(logits && cdf && workspace, "null pointer");
    NRRT_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0 (2D), 1 (3D) or 2 (raw heat map)");
    NRRT_REQUIRE(B >= 0 && cells > 0 && cells < ((int64_t)1 << 31), "bad sizes");
    if (ws_bytes < nrrt_cdf_workspace_bytes(B, cells)) return nrrt_fail(NRRT_ERR_WORKSPACE, "cdf workspace too small: %zu < %zu", ws_bytes, nrrt_cdf_workspace_bytes(B, cells));
    if (B == 0) return NRRT_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

    Tree t;
    t.build(0, cells, 0);
    std::vector<LeafOp> ops;
    std::vector<int32_t> level_off;
    level_off.push_back(0);
    for (int d = (int)t.by_depth.size() - 1; d >= 0; --d) {
        ops.insert(ops.end(), t.by_depth[d].begin(), t.by_depth[d].end());
        level_off.push_back((int32_t)ops.size());
    }
    NRRT_REQUIRE(level_off.size() <= 256, "pairwise tree too deep");
    const size_t max_leaves = (size_t)(cells / 64 + 2);
    NRRT_REQUIRE(t.leaves.size() <= max_leaves && ops.size() <= max_leaves, "pairwise tree larger than workspace");

    unsigned char* ws = static_cast<unsigned char*>(workspace);
    float4* stats = reinterpret_cast<float4*>(ws);
    ws += align_up((size_t)B * sizeof(float4), 256);
    LeafOp* d_leaves = reinterpret_cast<LeafOp*>(ws);
    ws += align_up(max_leaves * sizeof(LeafOp), 256);
    LeafOp* d_ops = reinterpret_cast<LeafOp*>(ws);
    ws += align_up(max_leaves * sizeof(LeafOp), 256);
    int32_t* d_level = reinterpret_cast<int32_t*>(ws);
    // pageable sources: cudaMemcpyAsync stages them before returning, so the vectors may die at scope exit
    NRRT_CUDA_TRY(cudaMemcpyAsync(d_leaves, t.leaves.data(), t.leaves.size() * sizeof(LeafOp), cudaMemcpyHostToDevice, st));
    if (!ops.empty()) NRRT_CUDA_TRY(cudaMemcpyAsync(d_ops, ops.data(), ops.size() * sizeof(LeafOp), cudaMemcpyHostToDevice, st));
    NRRT_CUDA_TRY(cudaMemcpyAsync(d_level, level_off.data(), level_off.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));

    const size_t smem = t.leaves.size() * sizeof(float);
    NRRT_REQUIRE(smem <= 200 * 1024, "map too large for the leaf-sum buffer");
    int dev = 0, sms = 0;
    NRRT_CUDA_TRY(cudaGetDevice(&dev));
    NRRT_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // streaming (bulk-copy pipelined) kernel when the rows satisfy the copy alignment; chunks = runs of whole leaves
    const size_t stream_smem = (size_t)kStages * kChunk * sizeof(float) + smem;
    if (cells % 4 == 0 && ((uintptr_t)logits & 15) == 0 && ((uintptr_t)cdf & 15) == 0 && stream_smem <= 220 * 1024) {
        std::vector<int32_t> chunk_leaf;
        int64_t chunk_begin = 0;
        chunk_leaf.push_back(0);
        for (size_t l = 0; l < t.leaves.size(); ++l) {
            const int64_t leaf_end = (int64_t)t.leaves[l].a + t.leaves[l].b;
            if (leaf_end - chunk_begin > kChunk) {  // this leaf opens a new chunk
                chunk_leaf.push_back((int32_t)l);
                chunk_begin = t.leaves[l].a;
            }
        }
        bool aligned = true;  // every chunk start must be 16-byte aligned (leaf offsets are multiples of 8 except for tiny rows)
        for (int32_t l : chunk_leaf) aligned = aligned && (t.leaves[l].a % 4 == 0);
        if (aligned && chunk_leaf.size() <= max_leaves) {
            int32_t* d_chunks = reinterpret_cast<int32_t*>(ws + 256 * sizeof(int32_t));
            NRRT_CUDA_TRY(cudaMemcpyAsync(d_chunks, chunk_leaf.data(), chunk_leaf.size() * sizeof(int32_t), cudaMemcpyHostToDevice, st));
            NRRT_CUDA_TRY(cudaFuncSetAttribute(cdf_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stream_smem));
            cdf_stream_kernel<<<B < sms ? B : sms, kSC + 32, stream_smem, st>>>(logits, mode, B, cells, d_leaves, (int)t.leaves.size(),
                                                                             d_ops, d_level, (int)level_off.size() - 1, d_chunks,
                                                                             (int)chunk_leaf.size(), stats, cdf);
            NRRT_CUDA_TRY(cudaGetLastError());
            return NRRT_OK;
        }
    }
    NRRT_CUDA_TRY(cudaFuncSetAttribute(cdf_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cdf_fused_kernel<<<B < sms ? B : sms, kFT, smem, st>>>(logits, mode, B, cells, d_leaves, (int)t.leaves.size(), d_ops, d_level,
                                                         (int)level_off.size() - 1, stats, cdf);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_draw_samples(const NrrtPlanParams* p, const uint32_t* occ_bits, const int32_t* task_to_map,
                                 const float* cdf, const int32_t* task_ids, int B, int16_t* samples, int8_t* neural_flag,
                                 int64_t* n_draws, int32_t* draw_status, void* stream) {
    NRRT_REQUIRE(p, "null params");
    NRRT_REQUIRE(p->dim == 2 || p->dim == 3, "dim must be 2 or 3, got %d", p->dim);
    NRRT_REQUIRE(B >= 0 && p->max_iterations >= 1 && p->max_reject >= 1, "bad sizes");
    if (B == 0) return NRRT_OK;
    NRRT_REQUIRE(occ_bits && task_to_map && task_ids && samples && draw_status, "null pointer");
    NRRT_REQUIRE(p->sampler_mode >= 0 && p->sampler_mode <= 2, "sampler_mode must be 0, 1 or 2");
    NRRT_REQUIRE(p->sampler_mode != 1 || cdf, "sampler_mode 1 (generate_neural_sample) needs a CDF");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DrawArgs a;
    a.p = *p; a.occ_bits = occ_bits; a.task_to_map = task_to_map; a.cdf = cdf; a.task_ids = task_ids; a.B = B;
    a.samples = samples; a.neural_flag = neural_flag; a.n_draws = n_draws; a.draw_status = draw_status;
    a.cells = (int64_t)p->shape[0] * p->shape[1] * (p->dim == 3 ? p->shape[2] : 1);
    a.words_per_map = (int)nrrt_words_per_map(a.cells);
    const int warps = 4;
    if (p->dim == 2) draw_kernel<2><<<(B + warps - 1) / warps, warps * 32, 0, st>>>(a);
    else draw_kernel<3><<<(B + warps - 1) / warps, warps * 32, 0, st>>>(a);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}
