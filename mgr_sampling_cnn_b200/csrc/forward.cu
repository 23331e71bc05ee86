// forward.cu — the whole CNN forward as ONE C-ABI call: nrrt_conv_forward_2d / nrrt_conv_forward_3d (include/nrrt.h).
//
// Host-side layer graph of UNet_cooler.forward (train.py:171-203) and ThreeD_UNet_cooler.forward (ThreeD_train.py:169-204)
// over the kernels of this library (nrrt_conv_layer = tcgen05 implicit GEMM, nrrt_stem_conv_maps, nrrt_coords_*,
// nrrt_poly*), with the exact rewrites of DESIGN.md 4.0:
//   * the encoder (conv1..conv5) and the encoder-side partial sums of the decoder's first convs run once per distinct
//     map, 8 maps per pass; the decoder runs once per task, in micro-batches (40 tasks in 2D, 10 in 3D);
//   * the interpolated coordinate channels never exist: their contribution is a per-task polynomial (2D: bilinear, 9
//     border cases; 3D: piecewise bilinear, 54 cases) evaluated in the consumers' epilogues or in a post-pass;
//   * 2D: conv6[0] is hoisted to the maps completely (16-case field nrrt_poly16 / nrrt_poly_relu), upconv7 -> conv7[0] and
//     upconv8 -> conv8[0] run as folded layers where the low-resolution grid holds whole 16x16 tiles;
//   * every torch.cat of the reference is a channel slice, a K split or a residual: no concatenated buffer is written.
// Nothing is allocated here: activations, coordinate-branch tensors and the two small index arrays are carved out of the
// caller's workspace.  The same graph is what mgr_sampling_cnn_b200/cnn.py drives layer by layer (kept for the unit tests
// of single layers); the parity tests compare both with the reference's logits.
#include <cuda_fp16.h>

#include <algorithm>
#include <vector>

#include "nrrt_common.cuh"

namespace {

constexpr int kEncBatch = 8;     // distinct maps per encoder pass
// tasks per coordinate-branch pass.  The branch is four tiny convs per task (4 tasks per CTA): a pass over 512 tasks is a
// 128-CTA launch that cannot fill 148 SMs (3.4 ms per pass in ncu, 27 ms per 4096-task step); 4096 tasks per pass
// cost 0.8 MB of workspace per task (polynomials + GEMM scratch) and run at full width.
constexpr int kPrepChunk2 = 4096;
constexpr int kPrepChunk3 = 128; // 3D (54-case field: 774 KB per task + 0.9 MB GEMM scratch)

struct Grid { int d, h, w; };
inline long long cells(const Grid& g) { return (long long)g.d * g.h * g.w; }

struct Act {  // channels-last fp16 activation [n][d][h][w][c]
    __half* p = nullptr;
    int n = 0, c = 0;
};

struct Bump {  // workspace carver (256-byte aligned pieces); with base == nullptr it only measures
    uint8_t* base;
    size_t off = 0;
    explicit Bump(void* b) : base(static_cast<uint8_t*>(b)) {}
    void* take(size_t bytes) {
        off = (off + 255) & ~(size_t)255;
        void* r = base ? base + off : nullptr;
        off += bytes;
        return r;
    }
    Act act(int n, const Grid& g, int c) {
        Act a;
        a.p = static_cast<__half*>(take((size_t)n * cells(g) * c * sizeof(__half)));
        a.n = n; a.c = c;
        return a;
    }
    float* f32(size_t count) { return static_cast<float*>(take(count * sizeof(float))); }
};

struct Opt {  // optional operands of one layer launch (mirrors cnn.py _Layer.run)
    const Act* res = nullptr;
    const float* poly = nullptr;
    int poly_cases = 0;
    const float* poly3 = nullptr;
    const Act* x2 = nullptr;
    const int32_t* idx = nullptr;
    const int32_t* idx2 = nullptr;
    const int32_t* res_idx = nullptr;
    int res_group = 0;
    float* logits = nullptr;
    const float* head_w = nullptr;
    float head_b = 0.f;
};

int run_layer(const NrrtConvW& L, int dims, void* st, const Act& x, const Grid& gin, int n, const Act* out, const Opt& o = Opt()) {
    NRRT_REQUIRE(L.weight, "layer (tag %d) is missing from the packed weights", L.tag);
    NrrtConvLayer l = {};
    l.dims = dims; l.kind = L.kind; l.tag = L.tag;
    l.n = n; l.in_d = gin.d; l.in_h = gin.h; l.in_w = gin.w;
    l.cin = L.cin; l.cout = L.cout;
    l.in = x.p; l.in_cstride = x.c; l.in_n = x.n;
    l.in_index = o.idx;
    if (o.x2) { l.in2 = o.x2->p; l.in2_cstride = o.x2->c; l.cin2 = L.cin2; l.in2_n = o.x2->n; l.in2_index = o.idx2; }
    l.weight = L.weight; l.scale = L.scale; l.shift = L.shift;
    l.post_scale = L.post_scale; l.post_shift = L.post_shift;
    if (out) { l.out = out->p; l.out_cstride = out->c; l.out_coffset = 0; }
    else { l.out_cstride = L.cout; }
    if (o.logits) { l.logits = o.logits; l.head_weight = o.head_w; l.head_bias = o.head_b; }
    if (o.res) {
        l.res = o.res->p; l.res_cstride = o.res->c; l.res_coffset = 0; l.res_n = o.res->n;
        l.res_index = o.res_idx; l.res_group = o.res_idx ? o.res_group : 0;
    }
    l.relu = (L.relu || o.poly3) ? 1 : 0;  // the fused 3D field carries the layer's ReLU with it
    l.poly = o.poly; l.poly_cases = o.poly_cases; l.poly3 = o.poly3;
    return nrrt_conv_layer(&l, st);
}

#define TRY(call)                 \
    do {                          \
        const int _rc = (call);   \
        if (_rc != NRRT_OK) return _rc; \
    } while (0)

struct Forward {
    const NrrtCnnWeights& w;
    int dims;
    Grid g[4];         // the four resolution levels (network axes d, h, w)
    bool fold;         // 2D: upconv7/conv7[0] and upconv8/conv8[0] as folded layers
    bool fuse3;        // 3D: conv8[0] adds its coordinate field in the epilogue
    int mb, pc;        // decoder micro-batch, coordinate-branch chunk
    void* st;
    // workspace pieces
    int32_t *d_maps = nullptr, *d_rel = nullptr;
    Act t0, x1, a0, b0, e2, a1, b1, d1, e3, a2, b2, d2, e4, a3, b3, d3, e5, r6, r7, r8, t6m;   // encoder set (kEncBatch maps)
    Act j, g7, i, g6, h, t8, t7, t6, c5;                                                        // decoder set (mb tasks)
    float* feat[4] = {nullptr, nullptr, nullptr, nullptr};   // coordinate features [pc][P][64/128/256/512]
    float *e_up6 = nullptr, *e_c6 = nullptr, *e_c7 = nullptr, *e_c8 = nullptr, *p16 = nullptr, *ef7 = nullptr, *ef8 = nullptr;
    float *a3d = nullptr, *work = nullptr;
    size_t work_floats = 0;

    Forward(const NrrtCnnWeights& w_, int d, int h, int ww, void* stream) : w(w_), dims(w_.dims), st(stream) {
        for (int lv = 0; lv < 4; ++lv)
            g[lv] = Grid{dims == 3 ? std::max(d >> lv, 1) : 1, std::max(h >> lv, 1), std::max(ww >> lv, 1)};
        fold = dims == 2 && std::min(h, ww) >= 64;
        fuse3 = dims == 3 && g[0].h % 16 == 0 && g[0].w % 16 == 0;
        mb = dims == 2 ? 40 : 10;
        pc = dims == 2 ? kPrepChunk2 : kPrepChunk3;
    }

    // carve (or, with a measuring Bump, size) the workspace for up to n_maps_pass maps per encoder pass and n_tasks tasks
    void carve(Bump& b, int n_tasks) {
        const int Be = kEncBatch, B = std::min(mb, std::max(n_tasks, 1)), P = std::min(pc, std::max(n_tasks, 1));
        d_maps = static_cast<int32_t*>(b.take(sizeof(int32_t) * Be));
        d_rel = static_cast<int32_t*>(b.take(sizeof(int32_t) * std::max(n_tasks, 1)));
        t0 = b.act(Be, g[0], 64); x1 = b.act(Be, g[0], 64); a0 = b.act(Be, g[0], 64); b0 = b.act(Be, g[0], 64); e2 = b.act(Be, g[0], 64);
        a1 = b.act(Be, g[1], 128); b1 = b.act(Be, g[1], 128); d1 = b.act(Be, g[1], 128); e3 = b.act(Be, g[1], 128);
        a2 = b.act(Be, g[2], 256); b2 = b.act(Be, g[2], 256); d2 = b.act(Be, g[2], 256); e4 = b.act(Be, g[2], 256);
        a3 = b.act(Be, g[3], 512); b3 = b.act(Be, g[3], 512); d3 = b.act(Be, g[3], 512); e5 = b.act(Be, g[3], 512);
        r6 = b.act(Be, g[2], 512); r7 = b.act(Be, g[1], 256); r8 = b.act(Be, g[0], 128);
        if (dims == 2) t6m = b.act(Be, g[2], 512);
        j = b.act(B, g[0], 128); g7 = b.act(B, g[1], 256); i = b.act(B, g[1], 128); g6 = b.act(B, g[2], 512); h = b.act(B, g[2], 256);
        if (!fold) { t8 = b.act(B, g[0], 64); t7 = b.act(B, g[1], 128); }
        if (dims == 3) { t6 = b.act(B, g[2], 512); c5 = b.act(B, g[3], 512); }
        const int pts = dims == 2 ? 4 : 6;
        const int fc[4] = {64, 128, 256, 512};
        for (int k = 0; k < 4; ++k) feat[k] = b.f32((size_t)P * pts * fc[k]);
        if (dims == 2) {
            e_up6 = b.f32((size_t)P * 4 * 4 * 512);
            e_c6 = b.f32((size_t)P * 9 * 4 * 512);
            e_c7 = b.f32((size_t)P * 9 * 4 * 256);
            e_c8 = b.f32((size_t)P * 9 * 4 * 128);
            p16 = b.f32((size_t)P * 16 * 4 * 512);
            if (fold) { ef7 = b.f32((size_t)P * 4 * 9 * 4 * 256); ef8 = b.f32((size_t)P * 4 * 9 * 4 * 128); }
            // scratch of nrrt_coords_poly_gemm (rows * 4 * (Cc + cases * n_uv * Cout)) and nrrt_poly16 (rows * 16 * 9 * Cout)
            work_floats = std::max({(size_t)P * 4 * (512 + 4 * 1 * 512), (size_t)P * 4 * (256 + 9 * 4 * 512),
                                    (size_t)P * 4 * (128 + 9 * 4 * 256), (size_t)P * 4 * (64 + 9 * 4 * 128), (size_t)P * 16 * 9 * 512});
        } else {
            e_c6 = b.f32((size_t)P * 54 * 4 * 512);
            e_c7 = b.f32((size_t)P * 54 * 4 * 256);
            e_c8 = b.f32((size_t)P * 54 * 4 * 128);
            a3d = b.f32((size_t)P * 4 * 2 * 256);
            work_floats = (size_t)P * 4 * (2 * 256 + 54 * 4 * 512);
        }
        work = b.f32(work_floats);
    }

    // ---- encoder + encoder-side partial sums for Be maps (rows d_maps[0..Be) of occ) ----
    int stage(const NrrtConvW (&blk)[2][3], const Act& xin, const Grid& gin, const Grid& gout, int n, Act& a, Act& b, Act* dsb, Act& fin) {
        const Act* res = &xin;
        if (blk[0][2].weight) {
            TRY(run_layer(blk[0][2], dims, st, xin, gin, n, dsb));
            res = dsb;
        }
        TRY(run_layer(blk[0][0], dims, st, xin, gin, n, &a));
        Opt o1; o1.res = res;
        TRY(run_layer(blk[0][1], dims, st, a, gout, n, &b, o1));
        TRY(run_layer(blk[1][0], dims, st, b, gout, n, &a));
        Opt o2; o2.res = &b;
        TRY(run_layer(blk[1][1], dims, st, a, gout, n, &fin, o2));
        return NRRT_OK;
    }

    int encode(const uint8_t* occ, int Be) {
        TRY(nrrt_stem_conv_maps(dims, occ, d_maps, Be, w.stem_cin, g[0].d, g[0].h, g[0].w, w.stem_w, w.stem_b, w.stem_cout, 64, t0.p, st));
        TRY(run_layer(w.stem2, dims, st, t0, g[0], Be, &x1));
        TRY(stage(w.enc[0], x1, g[0], g[0], Be, a0, b0, nullptr, e2));
        TRY(stage(w.enc[1], e2, g[0], g[1], Be, a1, b1, &d1, e3));
        TRY(stage(w.enc[2], e3, g[1], g[2], Be, a2, b2, &d2, e4));
        TRY(stage(w.enc[3], e4, g[2], g[3], Be, a3, b3, &d3, e5));
        if (dims == 2) {  // conv6[0] is linear in (upconv6(x5) | x4 | coords): all of its GEMM runs here, once per map
            TRY(run_layer(w.up6, dims, st, e5, g[3], Be, &t6m));
            Opt o; o.x2 = &e4;
            TRY(run_layer(w.c6full, dims, st, t6m, g[2], Be, &r6, o));
        } else {
            TRY(run_layer(w.c6e, dims, st, e4, g[2], Be, &r6));
        }
        TRY(run_layer(w.c7e, dims, st, e3, g[1], Be, &r7));
        TRY(run_layer(w.c8e, dims, st, e2, g[0], Be, &r8));
        return NRRT_OK;
    }

    // ---- coordinate branch for P tasks (train.py:179-190): features and the consumers' polynomials ----
    int prepare(const float* coords, int P) {
        float* const feats[4] = {feat[0], feat[1], feat[2], feat[3]};
        TRY(nrrt_coords_branch(coords, P, 2, dims == 2 ? 2 : 3, w.coords_w, w.coords_b, feats, st));
        auto inv = [](int v) { return v > 1 ? 1.0f / (float)(v - 1) : 0.0f; };
        if (dims == 2) {
            TRY(nrrt_coords_poly_gemm(feat[3], P, 512, w.mt_up6, 4, 1, 512, inv(g[3].h), inv(g[3].w), work, e_up6, st));
            TRY(nrrt_coords_poly_gemm(feat[2], P, 256, w.mt_c6, 9, 4, 512, inv(g[2].h), inv(g[2].w), work, e_c6, st));
            TRY(nrrt_coords_poly_gemm(feat[1], P, 128, w.mt_c7, 9, 4, 256, inv(g[1].h), inv(g[1].w), work, e_c7, st));
            TRY(nrrt_coords_poly_gemm(feat[0], P, 64, w.mt_c8, 9, 4, 128, inv(g[0].h), inv(g[0].w), work, e_c8, st));
            TRY(nrrt_poly16(e_up6, w.w6taps, e_c6, w.b6, P, 512, 512, g[3].h, g[3].w, work, p16, st));
            if (fold) {
                TRY(nrrt_poly_fold(e_c7, w.bc7, P, 256, g[2].h, g[2].w, ef7, st));
                TRY(nrrt_poly_fold(e_c8, w.bc8, P, 128, g[1].h, g[1].w, ef8, st));
            }
        } else {
            const float* mts[3] = {w.mt3_c8, w.mt3_c7, w.mt3_c6};
            float* es[3] = {e_c8, e_c7, e_c6};
            const int couts[3] = {128, 256, 512};
            for (int lv = 0; lv < 3; ++lv) {
                const int cc = 64 << lv;
                TRY(nrrt_coords_pieces_3d(feat[lv], P, cc, a3d, st));
                TRY(nrrt_coords_poly_gemm(a3d, P, 2 * cc, mts[lv], 54, 4, couts[lv], inv(g[lv].d), inv(g[lv].h), work, es[lv], st));
            }
        }
        return NRRT_OK;
    }

    // ---- decoder for B tasks; rel = device int32 [B] row of each task's map inside the current encoder pass;
    //      q0 = index of the first task inside the prepared coordinate chunk ----
    int decode(int B, const int32_t* rel, int q0, int res_group, float* logits) {
        Opt head; head.logits = logits; head.head_w = w.head_w; head.head_b = w.head_bias;
        if (dims == 2) {
            TRY(nrrt_poly_relu(r6.p, rel, p16 + (size_t)q0 * 16 * 4 * 512, g6.p, B, g[2].h, g[2].w, 512, st));
            TRY(run_layer(w.c61, dims, st, g6, g[2], B, &h));
            Opt o7; o7.res = &r7; o7.res_idx = rel; o7.res_group = res_group; o7.poly_cases = 9;
            Opt o8; o8.res = &r8; o8.res_idx = rel; o8.res_group = res_group; o8.poly_cases = 9;
            if (fold) {
                o7.poly = ef7 + (size_t)q0 * 4 * 9 * 4 * 256;
                TRY(run_layer(w.c7f, dims, st, h, g[2], B, &g7, o7));
                TRY(run_layer(w.c71, dims, st, g7, g[1], B, &i));
                o8.poly = ef8 + (size_t)q0 * 4 * 9 * 4 * 128;
                TRY(run_layer(w.c8f, dims, st, i, g[1], B, &j, o8));
            } else {
                TRY(run_layer(w.up7, dims, st, h, g[2], B, &t7));
                o7.poly = e_c7 + (size_t)q0 * 9 * 4 * 256;
                TRY(run_layer(w.c70, dims, st, t7, g[1], B, &g7, o7));
                TRY(run_layer(w.c71, dims, st, g7, g[1], B, &i));
                TRY(run_layer(w.up8, dims, st, i, g[1], B, &t8));
                o8.poly = e_c8 + (size_t)q0 * 9 * 4 * 128;
                TRY(run_layer(w.c80, dims, st, t8, g[0], B, &j, o8));
            }
            return run_layer(w.c81, dims, st, j, g[0], B, nullptr, head);
        }
        // 3D: only the bottleneck coordinate channels (second input of upconv6, (d/8, h/8, w/8) voxels) are materialised
        TRY(nrrt_coords_fill(feat[3] + (size_t)q0 * 6 * 512, B, 2, 3, 512, 3, g[3].d, g[3].h, g[3].w, c5.p, 512, 0, st));
        Opt u6; u6.x2 = &c5; u6.idx = rel;
        TRY(run_layer(w.up6, dims, st, e5, g[3], B, &t6, u6));
        Opt o6; o6.res = &r6; o6.res_idx = rel; o6.res_group = res_group;
        TRY(run_layer(w.c60, dims, st, t6, g[2], B, &g6, o6));
        TRY(nrrt_poly_relu_3d(g6.p, e_c6 + (size_t)q0 * 54 * 4 * 512, B, g[2].d, g[2].h, g[2].w, 512, st));
        TRY(run_layer(w.c61, dims, st, g6, g[2], B, &h));
        TRY(run_layer(w.up7, dims, st, h, g[2], B, &t7));
        Opt o7; o7.res = &r7; o7.res_idx = rel; o7.res_group = res_group;
        TRY(run_layer(w.c70, dims, st, t7, g[1], B, &g7, o7));
        TRY(nrrt_poly_relu_3d(g7.p, e_c7 + (size_t)q0 * 54 * 4 * 256, B, g[1].d, g[1].h, g[1].w, 256, st));
        TRY(run_layer(w.c71, dims, st, g7, g[1], B, &i));
        TRY(run_layer(w.up8, dims, st, i, g[1], B, &t8));
        Opt o8; o8.res = &r8; o8.res_idx = rel; o8.res_group = res_group;
        const float* e8 = e_c8 + (size_t)q0 * 54 * 4 * 128;
        if (fuse3) o8.poly3 = e8;
        TRY(run_layer(w.c80, dims, st, t8, g[0], B, &j, o8));
        if (!fuse3) TRY(nrrt_poly_relu_3d(j.p, e8, B, g[0].d, g[0].h, g[0].w, 128, st));
        return run_layer(w.c81, dims, st, j, g[0], B, nullptr, head);
    }
};

int check_shape(const NrrtCnnWeights* w, int dims, int d, int h, int ww) {
    NRRT_REQUIRE(w, "null weights");
    NRRT_REQUIRE(w->dims == dims, "these weights are for the %dD model", w->dims);
    NRRT_REQUIRE(h >= 8 && ww >= 8 && h % 8 == 0 && ww % 8 == 0 && (dims == 2 || (d >= 8 && d % 8 == 0)),
                 "spatial sizes must be multiples of 8 (three stride-2 stages)");
    return NRRT_OK;
}

int forward_impl(const NrrtCnnWeights* w, int dims, const uint8_t* occ, const int32_t* map_ids, int n_maps, const int32_t* t2r,
                 const float* coords, int n_tasks, int d, int h, int ww, int tasks_per_map, float* logits, void* ws,
                 size_t ws_bytes, void* stream) {
    TRY(check_shape(w, dims, d, h, ww));
    NRRT_REQUIRE(n_maps >= 0 && n_tasks >= 0, "negative sizes");
    if (n_tasks == 0) return NRRT_OK;
    NRRT_REQUIRE(occ && map_ids && t2r && coords && logits && ws && n_maps >= 1, "null pointer");
    for (int t = 0; t < n_tasks; ++t)
        NRRT_REQUIRE(t2r[t] >= 0 && t2r[t] < n_maps && (t == 0 || t2r[t] >= t2r[t - 1]),
                     "task_to_row must be non-decreasing with values in [0, n_maps) (task %d)", t);
    // 3D: the network's (d, h, w) axes are the map file's (z, x, y): nrrt_stem_conv_maps takes the network grid
    Forward f(*w, d, h, ww, stream);
    Bump measure(nullptr);
    f.carve(measure, n_tasks);
    if (measure.off > ws_bytes)
        return nrrt_fail(NRRT_ERR_WORKSPACE, "workspace too small: %zu bytes given, %zu needed", ws_bytes, measure.off);
    NRRT_REQUIRE(((uintptr_t)ws & 255) == 0, "workspace must be 256-byte aligned");
    Bump b(ws);
    f.carve(b, n_tasks);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    std::vector<int32_t> rel(n_tasks);
    for (int t = 0; t < n_tasks; ++t) rel[t] = t2r[t] % kEncBatch;
    NRRT_CUDA_TRY(cudaMemcpyAsync(f.d_rel, rel.data(), sizeof(int32_t) * n_tasks, cudaMemcpyHostToDevice, st));
    const long long n_cells = (long long)(dims == 3 ? d : 1) * h * ww;
    const int pts = dims == 2 ? 4 : 6;
    int prep_lo = 0, prep_hi = 0;  // tasks [prep_lo, prep_hi) have their coordinate branch in the workspace
    int lo = 0;
    for (int g0 = 0; g0 < n_maps; g0 += kEncBatch) {
        const int Be = std::min(kEncBatch, n_maps - g0);
        int hi = lo;
        while (hi < n_tasks && t2r[hi] < g0 + Be) ++hi;
        if (hi == lo) continue;  // none of these maps has a task
        NRRT_CUDA_TRY(cudaMemcpyAsync(f.d_maps, map_ids + g0, sizeof(int32_t) * Be, cudaMemcpyHostToDevice, st));
        TRY(f.encode(occ, Be));
        for (int i = lo; i < hi; i += f.mb) {
            const int j = std::min(hi, i + f.mb);
            if (j > prep_hi) {
                prep_lo = i;
                prep_hi = std::min(n_tasks, i + f.pc);
                TRY(f.prepare(coords + (size_t)prep_lo * 2 * (pts / 2), prep_hi - prep_lo));
            }
            // scheduling hint: consecutive tasks per map, when every map of the micro-batch has that many
            int rg = tasks_per_map > 1 ? tasks_per_map : 0;
            if (rg) {
                int run = 0;
                for (int t = i; t < j && rg; ++t) {
                    ++run;
                    if (t + 1 == j || t2r[t + 1] != t2r[t]) { if (run != tasks_per_map) rg = 0; run = 0; }
                }
            }
            TRY(f.decode(j - i, f.d_rel + i, i - prep_lo, rg, logits + (size_t)i * n_cells));
        }
        lo = hi;
    }
    return NRRT_OK;
}

}  // namespace

extern "C" size_t nrrt_conv_forward_workspace_bytes(const NrrtCnnWeights* w, int d, int h, int w_, int n_tasks) {
    if (!w || (w->dims != 2 && w->dims != 3) || h < 8 || w_ < 8 || n_tasks < 0) return 0;
    Forward f(*w, d, h, w_, nullptr);
    Bump measure(nullptr);
    f.carve(measure, n_tasks);
    return (measure.off + 255) & ~(size_t)255;
}

extern "C" int nrrt_conv_forward_2d(const NrrtCnnWeights* w, const uint8_t* occ_maps, const int32_t* map_ids, int n_maps,
                                    const int32_t* task_to_row, const float* coords, int n_tasks, int h, int w_,
                                    int tasks_per_map, float* logits, void* workspace, size_t ws_bytes, void* stream) {
    return forward_impl(w, 2, occ_maps, map_ids, n_maps, task_to_row, coords, n_tasks, 1, h, w_, tasks_per_map, logits, workspace,
                        ws_bytes, stream);
}

extern "C" int nrrt_conv_forward_3d(const NrrtCnnWeights* w, const uint8_t* occ_maps, const int32_t* map_ids, int n_maps,
                                    const int32_t* task_to_row, const float* coords, int n_tasks, int d, int h, int w_,
                                    int tasks_per_map, float* logits, void* workspace, size_t ws_bytes, void* stream) {
    return forward_impl(w, 3, occ_maps, map_ids, n_maps, task_to_row, coords, n_tasks, d, h, w_, tasks_per_map, logits, workspace,
                        ws_bytes, stream);
}
