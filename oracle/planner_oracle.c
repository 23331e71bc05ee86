/*
 * TEST INFRASTRUCTURE — CPU oracle for the planner path.  NOT part of the product: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library (oracle/planner_oracle.py wraps it).
 *
 * Scalar C restatement of the reference's RRT* planners, following SURVEY.md Appendix A line by line:
 *   rrt_star.py:63-215, neural_rrt.py:66-249, ThreeD_rrt_star.py:64-224, ThreeD_neural_rrt.py:70-265   (/root/reference)
 * Every function cites the reference lines it follows.  Arithmetic is IEEE binary64 with contraction OFF
 * (-ffp-contract=off) except the single FMA chain inside np.linalg.norm (OpenBLAS ddot).
 *
 * Parity pin: tests/test_oracle_vs_golden.py checks this file against traces recorded from the unmodified reference
 * (oracle/gen_golden.py -> tests/golden/*.npz): every node position, cost, parent index, iteration count and the ordered
 * collision-verdict list are compared bit-for-bit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NRRTO_MAXD 3

/* ---- status codes (mirror include/nrrt.h) ---- */
#define ST_SOLVED 0
#define ST_MAX_ITER_BEST_EFFORT 1
#define ST_NO_NODE_CONNECTED 2
#define ST_SAMPLER_STUCK 3

/* ------------------------------------------------------------------------------------------------------------------
 * Philox-4x32-10 sample stream (definition in oracle/philox_stream.py; replaces CPython `random`).
 * ------------------------------------------------------------------------------------------------------------------ */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

double nrrto_uniform(uint64_t seed, uint32_t task, uint64_t k) {
    uint64_t blk = k >> 1;
    uint32_t c[4] = {(uint32_t)blk, (uint32_t)(blk >> 32), task, 0x4E525254u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint32_t a = (k & 1) ? c[2] : c[0], b = (k & 1) ? c[3] : c[1];
    return (double)(((uint64_t)(a >> 5) << 26) + (uint64_t)(b >> 6)) / 9007199254740992.0;
}

static int stream_randint(uint64_t seed, uint32_t task, uint64_t *k, int n) { /* randint(0, n-1) */
    double u = nrrto_uniform(seed, task, (*k)++);
    int v = (int)floor(u * (double)n);
    return v < n - 1 ? v : n - 1;
}

/* ------------------------------------------------------------------------------------------------------------------
 * A.8 neural sampler CDF: neural_rrt.py:73-86 / ThreeD_neural_rrt.py:82-94 (flatten, min, subtract, np.sum, divide,
 * np.cumsum), hoisted out of the per-draw path.  All fp32.
 * ------------------------------------------------------------------------------------------------------------------ */
static float pairwise_sum_f32(const float *a, long n) { /* NumPy's pairwise_sum for contiguous float32 (A.8) */
    if (n < 8) {
        float res = 0.f; /* numpy starts from -0.0 only for the identity-less case; 0.f + x is exact here */
        res = (n > 0) ? a[0] : 0.f;
        for (long i = 1; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        float r[8];
        for (int k = 0; k < 8; ++k) r[k] = a[k];
        long i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) r[k] += a[i + k];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_sum_f32(a, n2) + pairwise_sum_f32(a + n2, n - n2);
    }
}

float nrrto_pairwise_sum_f32(const float *a, long n) { return pairwise_sum_f32(a, n); }

/* h: heat map as the planner stores it (2D: clamp(logits,-3,1)+10, neural_rrt.py:64; 3D: masked-250,
 * ThreeD_neural_rrt.py:68), flattened row-major.  cdf[j] = fl32(cdf[j-1] + fl32(fl32(h[j]-min)/S)). */
void nrrto_build_cdf(const float *h, long n, float *cdf, float *out_min, float *out_sum) {
    float m = h[0];
    for (long j = 1; j < n; ++j) if (h[j] < m) m = h[j];
    float *w = (float *)malloc(sizeof(float) * (size_t)n);
    for (long j = 0; j < n; ++j) w[j] = h[j] - m;
    float S = pairwise_sum_f32(w, n);
    float c = 0.f;
    for (long j = 0; j < n; ++j) {
        float p = w[j] / S;
        c = (j == 0) ? p : (c + p);
        cdf[j] = c;
    }
    free(w);
    if (out_min) *out_min = m;
    if (out_sum) *out_sum = S;
}

/* np.searchsorted(cdf, u) side='left', compared in fp64, then clamped (neural_rrt.py:89-96). */
static long cdf_search(const float *cdf, long n, double u) {
    long lo = 0, hi = n; /* first j with cdf[j] >= u; NaN entries compare as "not < u" */
    while (lo < hi) {
        long mid = lo + ((hi - lo) >> 1);
        if ((double)cdf[mid] < u) lo = mid + 1; else hi = mid;
    }
    if (lo > n - 1) lo = n - 1;
    if (lo < 0) lo = 0;
    return lo;
}

long nrrto_cdf_search(const float *cdf, long n, double u) { return cdf_search(cdf, n, u); }

/* ------------------------------------------------------------------------------------------------------------------
 * A.2 per-iteration sample drawing.  occ_free[cell] = 1 iff the reference treats the cell as free
 * (2D `occ[y,x] != 0`, rrt_star.py:67;  3D `occ[x,y,z] == 0`, ThreeD_rrt_star.py:69).  shape = array shape (axis0,1,2).
 *   uniform 2D: x=randint(0,W-1), y=randint(0,H-1) -> (y,x)        rrt_star.py:63-68
 *   uniform 3D: x=randint(0,map_width-1), y=randint(0,map_height-1), z=randint(0,map_depth-1) -> (x,y,z), where
 *               map_height,map_width,map_depth = shape (sic: x is drawn from shape[1] but indexes axis 0)  ThreeD_rrt_star.py:64-70
 *   neural: random() < bias chooses the sampler (neural_rrt.py:225); 2D neural has no occupancy test (:73-99);
 *           3D neural redraws until the voxel is free (ThreeD_neural_rrt.py:80-108).
 * Returns status 0 or ST_SAMPLER_STUCK (bounded rejection; the reference would spin forever).
 * ------------------------------------------------------------------------------------------------------------------ */
int nrrto_draw_samples(int dim, const int *shape, const uint8_t *occ_free, const float *cdf /*NULL = uniform module*/,
                       double neural_bias, uint64_t seed, uint32_t task, int n_iter, int max_reject,
                       int32_t *samples /*[n_iter][dim]*/, int8_t *neural_flag /*[n_iter] or NULL*/, int64_t *n_draws) {
    uint64_t k = 0;
    long ncell = (long)shape[0] * shape[1] * (dim == 3 ? shape[2] : 1);
    for (int it = 0; it < n_iter; ++it) {
        int use_neural = 0;
        if (cdf) use_neural = nrrto_uniform(seed, task, k++) < neural_bias;
        int32_t *s = samples + (size_t)it * dim;
        if (neural_flag) neural_flag[it] = (int8_t)use_neural;
        int rej = 0;
        for (;;) {
            if (use_neural) {
                double u = 0.0 + (1.0 - 0.0) * nrrto_uniform(seed, task, k++); /* random.uniform(0,1) */
                long idx = cdf_search(cdf, ncell, u);
                if (dim == 2) { s[0] = (int32_t)(idx / shape[1]); s[1] = (int32_t)(idx % shape[1]); break; }
                long yz = (long)shape[1] * shape[2];
                s[0] = (int32_t)(idx / yz); s[1] = (int32_t)((idx % yz) / shape[2]); s[2] = (int32_t)(idx % shape[2]);
                if (occ_free[idx]) break;
            } else if (dim == 2) {
                int x = stream_randint(seed, task, &k, shape[1]);
                int y = stream_randint(seed, task, &k, shape[0]);
                if (occ_free[(long)y * shape[1] + x]) { s[0] = y; s[1] = x; break; }
            } else {
                int x = stream_randint(seed, task, &k, shape[1]); /* map_width  = shape[1] */
                int y = stream_randint(seed, task, &k, shape[0]); /* map_height = shape[0] */
                int z = stream_randint(seed, task, &k, shape[2]);
                int xi = x < shape[0] ? x : shape[0] - 1, yi = y < shape[1] ? y : shape[1] - 1; /* cubes: no-op */
                if (occ_free[((long)xi * shape[1] + yi) * shape[2] + z]) { s[0] = x; s[1] = y; s[2] = z; break; }
            }
            if (++rej >= max_reject) { *n_draws = (int64_t)k; return ST_SAMPLER_STUCK; }
        }
    }
    *n_draws = (int64_t)k;
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Geometry primitives
 * ------------------------------------------------------------------------------------------------------------------ */
/* A.3: scipy.spatial.distance.euclidean -> sqrt(((d0*d0 + d1*d1) + d2*d2)), unfused, left to right. */
static inline double euclid(int dim, const double *u, const double *v) {
    double d0 = u[0] - v[0], d1 = u[1] - v[1];
    double s = d0 * d0 + d1 * d1;
    if (dim == 3) { double d2 = u[2] - v[2]; s = s + d2 * d2; }
    return sqrt(s);
}

/* `x ** 2` on a Python/NumPy float is libm pow(x, 2.0) (A.4); pow_mode 1 = x*x (what the GPU computes). */
static inline double sq(double x, int pow_mode) { return pow_mode ? x * x : pow(x, 2.0); }

typedef struct {
    int dim, shape[3];
    const uint8_t *occ_free;
    int32_t *trace; int trace_cap, trace_len; /* records of 4 int32: iteration, kind(0 connect,1 loop A,2 loop B), node j, verdict */
    double *seg; int seg_cap;                 /* optional: p1,p2 of every traced call, [trace_cap][2*dim] */
    int clamped;
} ctx_t;

/* A.5: is_collision_free (rrt_star.py:108-123, ThreeD_rrt_star.py:115-132). */
static int collision_free(ctx_t *c, const double *p1, const double *p2) {
    int dim = c->dim;
    double dl[NRRTO_MAXD];
    for (int k = 0; k < dim; ++k) dl[k] = p1[k] - p2[k];
    double q = dl[0] * dl[0];           /* np.linalg.norm -> x.dot(x) -> OpenBLAS ddot: FMA chain (A.5) */
    q = fma(dl[1], dl[1], q);
    if (dim == 3) q = fma(dl[2], dl[2], q);
    double D = sqrt(q);
    int same = 1;
    for (int k = 0; k < dim; ++k) same &= (p1[k] == p2[k]);
    if (same) return 0;                 /* `if point1 == point2: return False` */
    double step = D / 99.0;             /* np.linspace(0, D, 100): step = D/div, y = arange*step, y[-1] = stop */
    long prev = -1;
    for (int i = 0; i < 100; ++i) {
        double t = (i == 99) ? D : (double)i * step;
        long idx = 0;
        for (int k = 0; k < dim; ++k) {
            double v = p1[k] - ((t * dl[k]) / D);
            long ci = (long)v;          /* int(): truncation toward zero */
            if (ci < 0) { ci = 0; c->clamped++; }
            if (ci > c->shape[k] - 1) { ci = c->shape[k] - 1; c->clamped++; }
            idx = idx * c->shape[k] + ci;
        }
        if (idx == prev) continue;      /* consecutive duplicate cell: verdict-preserving skip */
        prev = idx;
        if (!c->occ_free[idx]) return 0;
    }
    return 1;
}

static int traced_cf(ctx_t *c, int it, int kind, int j, const double *p1, const double *p2) {
    int r = collision_free(c, p1, p2);
    if (c->trace && c->trace_len < c->trace_cap) {
        int32_t *t = c->trace + 4 * (size_t)c->trace_len;
        t[0] = it; t[1] = kind; t[2] = j; t[3] = r;
        if (c->seg) {
            double *s = c->seg + (size_t)c->trace_len * 2 * c->dim;
            for (int k = 0; k < c->dim; ++k) { s[k] = p1[k]; s[c->dim + k] = p2[k]; }
        }
    }
    c->trace_len++;
    return r;
}

/* ------------------------------------------------------------------------------------------------------------------
 * The main loop: rrt_star.py:184-215 (== neural_rrt.py:215-249, ThreeD_*:193-224 / 233-265), with the sample for each
 * iteration supplied by the caller (pre-drawn, A.2) and the radius table r[i-1] = compute_search_radius at i (A.1).
 *
 * Outputs (SoA, insertion order = node index): pos[(max_iter+2)][dim], cost[], parent[] (-1 for the start node).
 * On success the goal-reaching node is stored at index n_nodes (position := goal, rrt_star.py:205) but NOT counted in
 * n_nodes, exactly like the reference's `nodes` list.
 * ------------------------------------------------------------------------------------------------------------------ */
int nrrto_plan(int dim, const int *shape, const uint8_t *occ_free, const double *radius, const double *samples,
               const double *start, const double *goal, int max_iter, double goal_thr, int pow_mode,
               double *pos, double *cost, int32_t *parent, int32_t *near_scratch,
               int32_t *out_n_nodes, int32_t *out_iters, int32_t *out_status, int32_t *out_goal_node,
               int32_t *out_best_node, double *out_best_dist,
               int32_t *trace, int trace_cap, int32_t *out_trace_len, double *seg, int32_t *out_clamped) {
    ctx_t c;
    c.dim = dim; c.shape[0] = shape[0]; c.shape[1] = shape[1]; c.shape[2] = dim == 3 ? shape[2] : 1;
    c.occ_free = occ_free; c.trace = trace; c.trace_cap = trace_cap; c.trace_len = 0; c.seg = seg; c.clamped = 0;

    int n = 1;                                    /* self.nodes = [self.start_node]  (rrt_star.py:57) */
    for (int k = 0; k < dim; ++k) pos[k] = start[k];
    cost[0] = 0.0; parent[0] = -1;
    double best = INFINITY; int best_node = -1, goal_node = -1, iters = 0;

    for (int i = 0; i < max_iter; ++i) {
        iters = i + 1;
        double r = radius[i];
        const double *smp = samples + (size_t)i * dim;

        /* find_nearest_neighbor (rrt_star.py:70-80): strict <, insertion order */
        int ni = -1; double md = INFINITY;
        for (int j = 0; j < n; ++j) {
            double d = euclid(dim, pos + (size_t)j * dim, smp);
            if (d < md) { md = d; ni = j; }
        }
        const double *from = pos + (size_t)ni * dim;

        /* steer (rrt_star.py:82-98; A.4) */
        double dr[NRRTO_MAXD], np_[NRRTO_MAXD];
        for (int k = 0; k < dim; ++k) dr[k] = smp[k] - from[k];
        double s = sq(dr[0], pow_mode) + sq(dr[1], pow_mode);
        if (dim == 3) s = s + sq(dr[2], pow_mode);
        double dist = sqrt(s);
        if (dist > r) for (int k = 0; k < dim; ++k) dr[k] = (dr[k] * r) / dist;
        s = sq(dr[0], pow_mode) + sq(dr[1], pow_mode);
        if (dim == 3) s = s + sq(dr[2], pow_mode);
        dist = sqrt(s);
        double ncost = cost[ni] + dist;
        for (int k = 0; k < dim; ++k) np_[k] = from[k] + dr[k];
        int npar = ni;

        /* can_connect_nodes (rrt_star.py:100-106) */
        if (!traced_cf(&c, iters, 0, ni, from, np_)) continue;

        /* rewire_tree (rrt_star.py:125-147): near set once, insertion order, <= r */
        int nn = 0;
        for (int j = 0; j < n; ++j)
            if (euclid(dim, np_, pos + (size_t)j * dim) <= r) near_scratch[nn++] = j;
        for (int q = 0; q < nn; ++q) {            /* loop A: choose parent */
            int j = near_scratch[q];
            double cj = cost[j] + euclid(dim, pos + (size_t)j * dim, np_);
            if (cj < ncost && traced_cf(&c, iters, 1, j, pos + (size_t)j * dim, np_)) { npar = j; ncost = cj; }
        }
        for (int q = 0; q < nn; ++q) {            /* loop B: rewire neighbours (costs NOT propagated) */
            int j = near_scratch[q];
            double cj = ncost + euclid(dim, np_, pos + (size_t)j * dim);
            if (cj < cost[j] && traced_cf(&c, iters, 2, j, np_, pos + (size_t)j * dim)) { parent[j] = n; cost[j] = cj; }
        }

        /* goal_reached (rrt_star.py:156-161) */
        double dg = euclid(dim, np_, goal);
        if (dg < best) { best = dg; best_node = n; }
        for (int k = 0; k < dim; ++k) pos[(size_t)n * dim + k] = np_[k];
        cost[n] = ncost; parent[n] = npar;
        if (dg <= goal_thr) {                     /* rrt_star.py:203-208: snap, break, not appended */
            goal_node = n;
            for (int k = 0; k < dim; ++k) pos[(size_t)n * dim + k] = goal[k];
            break;
        }
        n++;                                      /* self.nodes.append(new_node) */
    }
    *out_n_nodes = n; *out_iters = iters; *out_goal_node = goal_node; *out_best_node = best_node; *out_best_dist = best;
    *out_status = goal_node >= 0 ? ST_SOLVED : (best_node >= 0 ? ST_MAX_ITER_BEST_EFFORT : ST_NO_NODE_CONNECTED);
    if (out_trace_len) *out_trace_len = c.trace_len;
    if (out_clamped) *out_clamped = c.clamped;
    return 0;
}

/* find_path (rrt_star.py:163-172) + calculate_length (test_network_and_algorithms.py:52-56). Returns #points. */
int nrrto_path(int dim, const double *pos, const int32_t *parent, int end_node, int cap, int32_t *path_idx, double *length) {
    int m = 0;
    for (int j = end_node; j >= 0 && m < cap; j = parent[j]) path_idx[m++] = j;
    for (int a = 0, b = m - 1; a < b; ++a, --b) { int32_t t = path_idx[a]; path_idx[a] = path_idx[b]; path_idx[b] = t; }
    double L = 0.0;
    for (int q = 0; q + 1 < m; ++q) L += euclid(dim, pos + (size_t)path_idx[q] * dim, pos + (size_t)path_idx[q + 1] * dim);
    if (length) *length = L;
    return m;
}

/* ------------------------------------------------------------------------------------------------------------------
 * Batched driver used by bench.py's cpu_baseline / --impl reference leg.  Serial over its B tasks; the Python wrapper
 * fans task ranges out over host threads (ctypes releases the GIL; libgomp is not in this image).
 * occ_free: [n_maps][cells]; cdf: [B][cells] or NULL; starts/goals int32 [B][dim].
 * ------------------------------------------------------------------------------------------------------------------ */
int nrrto_plan_batch(int dim, const int *shape, const uint8_t *occ_free, const int32_t *task_to_map, const float *cdf,
                     double neural_bias, const double *radius, const int32_t *starts, const int32_t *goals,
                     uint64_t seed, const int32_t *task_ids, int B, int max_iter, double goal_thr, int max_reject, int pow_mode,
                     int32_t *n_nodes, int32_t *iters, int32_t *status, double *path_len) {
    long ncell = (long)shape[0] * shape[1] * (dim == 3 ? shape[2] : 1);
    for (int b = 0; b < B; ++b) {
        int32_t *smp_i = (int32_t *)malloc(sizeof(int32_t) * (size_t)max_iter * dim);
        double *smp = (double *)malloc(sizeof(double) * (size_t)max_iter * dim);
        double *pos = (double *)malloc(sizeof(double) * (size_t)(max_iter + 2) * dim);
        double *cost = (double *)malloc(sizeof(double) * (size_t)(max_iter + 2));
        int32_t *par = (int32_t *)malloc(sizeof(int32_t) * (size_t)(max_iter + 2));
        int32_t *scr = (int32_t *)malloc(sizeof(int32_t) * (size_t)(max_iter + 2));
        const uint8_t *occ = occ_free + (size_t)task_to_map[b] * ncell;
        int64_t nd;
        int st = nrrto_draw_samples(dim, shape, occ, cdf ? cdf + (size_t)b * ncell : NULL, neural_bias, seed,
                                    (uint32_t)task_ids[b], max_iter, max_reject, smp_i, NULL, &nd);
        if (st) { status[b] = st; n_nodes[b] = 1; iters[b] = 0; path_len[b] = 0.0; }
        else {
            for (long q = 0; q < (long)max_iter * dim; ++q) smp[q] = (double)smp_i[q];
            double s0[3], g0[3], bd; int32_t gn, bn, cl, tl;
            for (int k = 0; k < dim; ++k) { s0[k] = starts[b * dim + k]; g0[k] = goals[b * dim + k]; }
            nrrto_plan(dim, shape, occ, radius, smp, s0, g0, max_iter, goal_thr, pow_mode, pos, cost, par, scr,
                       &n_nodes[b], &iters[b], &status[b], &gn, &bn, &bd, NULL, 0, &tl, NULL, &cl);
            int end = gn >= 0 ? gn : bn;
            path_len[b] = 0.0;
            if (end >= 0) nrrto_path(dim, pos, par, end, max_iter + 2, scr, &path_len[b]);
        }
        free(smp_i); free(smp); free(pos); free(cost); free(par); free(scr);
    }
    return 0;
}
