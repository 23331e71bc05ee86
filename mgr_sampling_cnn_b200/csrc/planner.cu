// planner.cu — the RRT* main loop of the reference, one planning task per CTA (sm_100a).
//
// Replaces  rrt_star.py:70-215, neural_rrt.py:101-249, ThreeD_rrt_star.py:72-224, ThreeD_neural_rrt.py:112-265
// (find_nearest_neighbor, steer, can_connect_nodes, is_collision_free, rewire_tree, find_nearby_nodes, goal_reached,
// find_path, rrt_star) and calculate_length (test_network_and_algorithms.py:52-56).  Numeric contract: SURVEY.md
// Appendix A — fp64, every product/sum rounded separately except the FMA chain of np.linalg.norm in the collision
// test; nearest neighbour = lowest index among minimal sqrt'ed distances; near set uses `<=` on the sqrt'ed distance.
//
// Layout per CTA (dynamic shared memory):
//   packed node positions [cap] (node index = insertion order) | occupancy bit grid of the task's map | near bit-mask |
//   reduction slots.  Positions: almost every node IS its integer sample point (steer only clips when the sample is
//   farther than the search radius), so a node is one 32-bit (2D) / 64-bit (3D) word of int16 coordinates; a node
//   with a non-integer coordinate stores a sentinel and keeps its fp64 position in the caller's `pos` array (global
//   memory, read through L2).  cost[] and parent[] live in the caller's output arrays too: they are touched only for
//   the ~20 near nodes of an iteration.
//   Occupancy is what buys throughput (the kernel is a latency-bound chain of ~5000 dependent iterations per task,
//   profiles/planner_r1.md): round 1 staged the task's bit grid in shared memory (53 KiB per 2D task, 105 KiB per 3D task:
//   four / two tasks per SM, and 111 / 120 registers per thread capped it there as well).  Round 2: the production kernels
//   read the map bits from global memory through L1 (ld.global.nc; the tasks resident on an SM belong to one or two maps, so
//   their 32 / 64 KiB grids stay in L1 -- same speed as the staged copy at equal occupancy) and are compiled for 8 (2D, 64
//   registers) / 4 (3D, 64 registers) CTAs per SM: 21 / 41 KiB of shared memory per task.  Measured, 4096 2D uniform tasks:
//   209.7 ms (4 CTAs, staged) -> 207.1 (4, L1) -> 183.3 (5) -> 173.0 (6) -> 168.0 (7) -> 164.8 (8) -> 182.8 (10: spills take over);
//   1024 3D tasks: 126.8 -> 104.4 ms (profiles/planner_r2.md).
// Parallelism inside an iteration: the two O(n) scans run over all threads; every warp evaluates the connect test
// redundantly (no barrier); choose-parent/rewire candidates are tested speculatively, one candidate per warp at a time,
// 100 samples over 32 lanes.  The outcome of loop A does not depend on evaluation order (it is the lexicographic
// minimum of (cost, index) over the free candidates), so speculation is exact; TRACE mode replays the reference's
// sequential order on warp 0 to emit exactly the collision calls the reference makes.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "nrrt_common.cuh"

namespace {

constexpr int kWarp = 32;

struct PlanArgs {
    NrrtPlanParams p;
    const uint32_t* occ_bits;
    const int32_t* task_to_map;
    const double* radius;
    const int16_t* samples;
    const int32_t* starts;
    const int32_t* goals;
    const int32_t* draw_status;
    NrrtPlanOut out;
    int B;
    int32_t* queue;
    int words_per_map;
    int cap;         // node capacity in shared memory (>= max_iterations + 2, multiple of 4)
    int mask_words;  // words of the near mask
};

struct Slot {
    double v;
    int j;
    int pad;
};

// lexicographic "reference order" combine for the nearest-neighbour reduction: lower sqrt(s) wins, ties -> lower j.
// sqrt is only evaluated when the squares are within 2^-50 relative (squares alone mis-order sqrt ties).
__device__ __forceinline__ bool nn_a_wins(double sa, int ja, double sb, int jb) {
    if (sa == sb) return ja < jb;
    const double lo = fmin(sa, sb), hi = fmax(sa, sb);
    if (hi <= __dmul_rn(lo, NRRT_K_HI)) {
        const double da = sqrt(sa), db = sqrt(sb);
        return da < db || (da == db && ja < jb);
    }
    return sa < sb;
}

// Rare slow path of the nearest-neighbour scan (two squares within 2^-50 relative): compare like the reference, on the
// sqrt'ed values.  Kept out of line so the two DSQRT sequences are not if-converted into the hot loop.
__device__ __noinline__ bool nn_take_slow(double s, double bs) { return sqrt(s) < sqrt(bs); }

// Warp argmin in reference order.  Non-negative doubles order like their bit patterns, so the common case is an integer
// min + REDUX for the lowest index among exact ties; only when two different squares lie within 8 ulps (the window
// in which sqrt rounding can merge or reorder them) does the warp fall back to the sqrt-aware combine.
__device__ __forceinline__ void nn_warp_reduce(double& bs, int& bj) {
    const long long key = __double_as_longlong(bs);
    long long kmin = key;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
    const bool tie = key == kmin;
    if (!__any_sync(0xffffffffu, !tie && (key - kmin <= 8))) {
        bj = (int)__reduce_min_sync(0xffffffffu, tie ? (unsigned)bj : 0x7fffffffu);
        bs = __longlong_as_double(kmin);
        return;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double os = __shfl_xor_sync(0xffffffffu, bs, o);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
        if (!nn_a_wins(bs, bj, os, oj)) { bs = os; bj = oj; }
    }
}

// is_collision_free (rrt_star.py:108-123 / ThreeD_rrt_star.py:115-132), one warp, 100 samples over lanes.
// Grid bit = 1 means the reference treats the cell as free.
//
// The reference evaluates c_k = int(p1_k - ((t_i * d_k) / D)) with three separately rounded operations per coordinate;
// the fp64 division dominates the cost.  A sample's cell is decided by floor(), so the exact arithmetic only matters
// when the value is within rounding distance of an integer.  Fast path: v' = fma(-t_i, d_k / D, p1_k) differs from the
// reference value by < 1e-9 (a few ulps of a number <= 2^15), so whenever v' is farther than 1e-6 from every integer
// both truncate to the same cell; otherwise (segment ends on integer cells, axis-parallel moves, coincidences) the
// lane re-evaluates the reference's exact sequence.  t = 0 and d_k = 0 are exact without any division.

// ---- packed node positions ------------------------------------------------------------------------------------------
template <int DIM> struct PosWord { typedef uint32_t T; };
template <> struct PosWord<3> { typedef unsigned long long T; };
constexpr int kPosSentinel = -32768;  // coordinate 0 of a node whose fp64 position lives in global memory

template <int DIM>
__device__ __forceinline__ void load_pos(const typename PosWord<DIM>::T* __restrict__ pw, const double* gpos, int j,
                                         double (&o)[DIM]) {
    const typename PosWord<DIM>::T w = pw[j];
    const int k0 = (int)(int16_t)(uint16_t)(w & 0xffffu);
    if (k0 != kPosSentinel) {
        o[0] = (double)k0;
        o[1] = (double)(int)(int16_t)(uint16_t)((w >> 16) & 0xffffu);
        if (DIM == 3) o[DIM - 1] = (double)(int)(int16_t)(uint16_t)(((unsigned long long)w >> 32) & 0xffffu);
    } else {
#pragma unroll
        for (int k = 0; k < DIM; ++k) o[k] = __ldcg(gpos + (size_t)j * DIM + k);  // written by this CTA before a barrier
    }
}
// returns the packed word; *exact = false -> sentinel (the caller stores the fp64 position in global memory)
template <int DIM>
__device__ __forceinline__ typename PosWord<DIM>::T encode_pos(const double (&v)[DIM], bool* exact) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < DIM; ++k) ok = ok && v[k] == rint(v[k]) && v[k] > -32768.0 && v[k] < 32768.0;
    *exact = ok;
    if (!ok) return (typename PosWord<DIM>::T)(uint16_t)(int16_t)kPosSentinel;
    typename PosWord<DIM>::T w = 0;
#pragma unroll
    for (int k = 0; k < DIM; ++k) w |= (typename PosWord<DIM>::T)(uint16_t)(int16_t)__double2int_rn(v[k]) << (16 * k);
    return w;
}

template <int DIM>
struct Segment {  // per-segment constants of is_collision_free, identical in every lane
    double p1[DIM], dl[DIM], u[DIM], D, step;
    bool same;
};

template <int DIM>
__device__ __forceinline__ Segment<DIM> segment_setup(const double (&p1)[DIM], const double (&p2)[DIM]) {
    Segment<DIM> sg;
    sg.same = true;
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
        sg.p1[k] = p1[k];
        sg.dl[k] = __dsub_rn(p1[k], p2[k]);
        sg.same = sg.same && (p1[k] == p2[k]);
    }
    double q = __dmul_rn(sg.dl[0], sg.dl[0]);  // np.linalg.norm -> ddot: FMA chain
    q = __fma_rn(sg.dl[1], sg.dl[1], q);
    if (DIM == 3) q = __fma_rn(sg.dl[DIM - 1], sg.dl[DIM - 1], q);
    sg.D = sqrt(q);
    sg.step = __ddiv_rn(sg.D, 99.0);  // np.linspace(0, D, 100)
#pragma unroll
    for (int k = 0; k < DIM; ++k) sg.u[k] = __ddiv_rn(sg.dl[k], sg.D);
    return sg;
}

// sample i (0..99) of the segment: true iff its cell is an obstacle
// GS = true: `grid` is the CTA's shared-memory copy of the map; false: the map in global memory, read through L1
// (ld.global.nc: the bit grids of the few maps whose tasks are resident on an SM stay in its L1).
template <int DIM, bool GS = true>
__device__ __forceinline__ bool sample_blocked(const Segment<DIM>& sg, int i, const uint32_t* __restrict__ grid,
                                               const int (&shape)[3]) {
    const double t = (i == 99) ? sg.D : __dmul_rn((double)i, sg.step);
    int idx = 0;
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
        double v;
        if (i == 0 || sg.dl[k] == 0.0) {
            v = sg.p1[k];  // (0 * d) / D = +-0 and p1 - +-0 = p1: exact
        } else {
            v = __fma_rn(-t, sg.u[k], sg.p1[k]);
            if (fabs(v - rint(v)) <= 1e-6)  // rounding could change the cell: the reference's exact sequence
                v = __dsub_rn(sg.p1[k], __ddiv_rn(__dmul_rn(t, sg.dl[k]), sg.D));
        }
        int ci = __double2int_rz(v);  // int(): truncation toward zero
        ci = max(0, min(ci, shape[k] - 1));
        idx = idx * shape[k] + ci;
    }
    const uint32_t wd = GS ? grid[idx >> 5] : __ldg(grid + (idx >> 5));
    return ((wd >> (idx & 31)) & 1u) == 0u;
}

template <int DIM, bool GS = true>
__device__ __forceinline__ bool collision_free_warp(const double (&p1)[DIM], const double (&p2)[DIM],
                                                    const uint32_t* __restrict__ grid, const int (&shape)[3], int lane) {
    const Segment<DIM> sg = segment_setup<DIM>(p1, p2);
    if (sg.same) return false;  // `if point1 == point2: return False`
#pragma unroll 1
    for (int i = lane; i < 128; i += kWarp) {
        const bool blocked = i < 100 && sample_blocked<DIM, GS>(sg, i, grid, shape);
        if (__any_sync(0xffffffffu, blocked)) return false;
    }
    return true;
}

template <int DIM, int NT, bool TRACE, bool GS = true, int MINB = (DIM == 2 ? 4 : 2)>
__global__ void __launch_bounds__(NT, MINB) plan_kernel(const PlanArgs a) {
    constexpr int NW = NT / kWarp;
    typedef typename PosWord<DIM>::T PW;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PW* pw = reinterpret_cast<PW*>(smem_raw);                                // [cap] packed positions
    uint32_t* grid_s = reinterpret_cast<uint32_t*>(pw + a.cap);              // GS: [words_per_map] (padded to 4)
    uint32_t* mask = grid_s + (GS ? ((a.words_per_map + 3) & ~3) : 0);       // [mask_words]
    const uint32_t* grid = grid_s;
    Slot* slot_nn = reinterpret_cast<Slot*>(mask + ((a.mask_words + 3) & ~3)); // [2][NW]
    Slot* slot_a = slot_nn + 2 * NW;                                         // [NW]
    __shared__ int s_task;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int max_iter = a.p.max_iterations;
    const int shape[3] = {a.p.shape[0], a.p.shape[1], DIM == 3 ? a.p.shape[2] : 1};
    const double thr = a.p.goal_threshold;
    int cur_map = -1;

    for (;;) {
        __syncthreads();  // previous task fully retired (smem reuse, s_task reuse)
        if (tid == 0) s_task = atomicAdd(a.queue, 1);
        __syncthreads();
        const int task = s_task;
        if (task >= a.B) break;

        double* cost = a.out.cost + (size_t)task * a.p.node_stride;
        int32_t* parent = a.out.parent + (size_t)task * a.p.node_stride;
        double* gpos = a.out.pos + (size_t)task * a.p.node_stride * DIM;  // fp64 positions of the non-integer nodes (+ final dump)
        int32_t* trace = TRACE ? a.out.trace + (size_t)task * a.out.trace_cap * 4 : nullptr;
        int trace_n = 0;

        if (a.draw_status && a.draw_status[task] != 0) {  // sampler gave up: report and skip
            if (tid == 0) {
                a.out.n_nodes[task] = 1; a.out.iterations[task] = 0; a.out.status[task] = a.draw_status[task];
                a.out.goal_node[task] = -1; a.out.best_node[task] = -1; a.out.best_dist[task] = INFINITY;
                if (a.out.path_n) a.out.path_n[task] = 0;
                if (a.out.path_len) a.out.path_len[task] = 0.0;
                if (a.out.trace_n) a.out.trace_n[task] = 0;
            }
            continue;
        }

        // ---- stage the task's occupancy bits (16-byte loads; consecutive tasks of one map reuse the copy) ----
        const int map = a.task_to_map[task];
        if (!GS) {
            grid = a.occ_bits + (size_t)map * a.words_per_map;
        } else if (map != cur_map) {
            const uint4* src = reinterpret_cast<const uint4*>(a.occ_bits + (size_t)map * a.words_per_map);
            uint4* dst = reinterpret_cast<uint4*>(grid_s);
            for (int i = tid; i < (a.words_per_map >> 2); i += NT) dst[i] = __ldg(src + i);
            for (int i = (a.words_per_map & ~3) + tid; i < a.words_per_map; i += NT)
                grid_s[i] = __ldg(a.occ_bits + (size_t)map * a.words_per_map + i);
            cur_map = map;
        }
        double goal[DIM];
#pragma unroll
        for (int k = 0; k < DIM; ++k) goal[k] = (double)a.goals[task * DIM + k];
        if (tid == 0) {  // self.nodes = [Node(start)]
            double st[DIM];
#pragma unroll
            for (int k = 0; k < DIM; ++k) st[k] = (double)a.starts[task * DIM + k];
            bool exact;
            pw[0] = encode_pos<DIM>(st, &exact);
            if (!exact) {
#pragma unroll
                for (int k = 0; k < DIM; ++k) gpos[k] = st[k];
            }
            cost[0] = 0.0; parent[0] = -1;
        }
        __syncthreads();

        int n = 1, best_node = -1, goal_node = -1, iters = 0;
        double best = INFINITY;
        const int16_t* smp_base = a.samples + (size_t)task * max_iter * DIM;
        double r_next = __ldg(a.radius);
        double pend[DIM];   // newest node, kept in registers: its shared-memory copy becomes visible at barrier (1)
        PW pend_w = 0;      // ... and its packed word
        int pend_j = -1;
#pragma unroll
        for (int k = 0; k < DIM; ++k) pend[k] = 0.0;
        int16_t smp_next[DIM];
#pragma unroll
        for (int k = 0; k < DIM; ++k) smp_next[k] = __ldg(smp_base + k);

        for (int it = 0; it < max_iter; ++it) {
            iters = it + 1;
            const double r = r_next;
            double smp[DIM];
            int si[DIM];  // the sample is always an integer cell
#pragma unroll
            for (int k = 0; k < DIM; ++k) { si[k] = (int)smp_next[k]; smp[k] = (double)smp_next[k]; }
            if (it + 1 < max_iter) {  // next iteration's inputs travel while this one computes
                r_next = __ldg(a.radius + it + 1);
#pragma unroll
                for (int k = 0; k < DIM; ++k) smp_next[k] = __ldg(smp_base + (it + 1) * DIM + k);
            }

            // ---- find_nearest_neighbor: scan + argmin (strict <, lowest index wins) ----
            // Integer nodes against the integer sample: the squared distance is an exact integer, distinct integers below
            // 2^32 have distinct fp64 square roots, so integer order IS the reference's order of sqrt'ed distances (and
            // it is the value the fp64 path would compute).  Only the rare non-integer nodes take the fp64 path.
            double bs = INFINITY;
            int bj = 0x7fffffff;
            uint32_t bsi = 0xffffffffu;
            int bji = 0x7fffffff;
            for (int j = tid; j < n; j += NT) {
                const bool is_pend = j == pend_j;
                const PW w = is_pend ? pend_w : pw[j];
                const int k0 = (int)(int16_t)(uint16_t)(w & 0xffffu);
                if (k0 != kPosSentinel) {
                    const int d0 = k0 - si[0];
                    const int d1 = (int)(int16_t)(uint16_t)((w >> 16) & 0xffffu) - si[1];
                    uint32_t sq = (uint32_t)(d0 * d0) + (uint32_t)(d1 * d1);
                    if (DIM == 3) {
                        const int d2 = (int)(int16_t)(uint16_t)(((unsigned long long)w >> 32) & 0xffffu) - si[DIM - 1];
                        sq += (uint32_t)(d2 * d2);
                    }
                    if (sq < bsi) { bsi = sq; bji = j; }
                } else {
                    double pj[DIM];
#pragma unroll
                    for (int k = 0; k < DIM; ++k) pj[k] = is_pend ? pend[k] : __ldcg(gpos + (size_t)j * DIM + k);
                    const double s = nrrt_sqdist<DIM>(pj, smp);
                    if (s < bs) {
                        bool take = true;
                        if (s > __dmul_rn(bs, NRRT_K_LO)) take = nn_take_slow(s, bs);  // sqrt-tie hazard, rare
                        if (take) { bs = s; bj = j; }
                    }
                }
            }
            if (bji != 0x7fffffff) {  // merge the integer winner (exact as a double) with the fp64 one
                const double sd = (double)bsi;
                if (bj == 0x7fffffff || nn_a_wins(sd, bji, bs, bj)) { bs = sd; bj = bji; }
            }
            nn_warp_reduce(bs, bj);
            Slot* snn = slot_nn + (it & 1) * NW;
            if (lane == 0) { snn[warp].v = bs; snn[warp].j = bj; }
            __syncthreads();  // (1)
            {
                // all threads combine the NW warp results redundantly (no second barrier): integer keys first
                long long kmin = __double_as_longlong(snn[0].v);
#pragma unroll
                for (int w = 1; w < NW; ++w) kmin = min(kmin, __double_as_longlong(snn[w].v));
                bool hazard = false;
                int jmin = 0x7fffffff;
#pragma unroll
                for (int w = 0; w < NW; ++w) {
                    const long long kw = __double_as_longlong(snn[w].v);
                    if (kw == kmin) jmin = min(jmin, snn[w].j);
                    else hazard = hazard || (kw - kmin <= 8);
                }
                if (!hazard) {
                    bs = __longlong_as_double(kmin); bj = jmin;
                } else {  // two different squares within a few ulps: compare the sqrt'ed values like the reference
                    bs = snn[0].v; bj = snn[0].j;
                    for (int w = 1; w < NW; ++w) {
                        const double os = snn[w].v;
                        const int oj = snn[w].j;
                        if (!nn_a_wins(bs, bj, os, oj)) { bs = os; bj = oj; }
                    }
                }
            }
            const int ni = bj;

            // ---- steer (rrt_star.py:82-98; A.4 with x*x for x**2) ----
            double from[DIM], dr[DIM], npos[DIM];
            load_pos<DIM>(pw, gpos, ni, from);
#pragma unroll
            for (int k = 0; k < DIM; ++k) dr[k] = __dsub_rn(smp[k], from[k]);
            double s2 = __dadd_rn(__dmul_rn(dr[0], dr[0]), __dmul_rn(dr[1], dr[1]));
            if (DIM == 3) s2 = __dadd_rn(s2, __dmul_rn(dr[DIM - 1], dr[DIM - 1]));
            double dist = sqrt(s2);
            if (dist > r) {
#pragma unroll
                for (int k = 0; k < DIM; ++k) dr[k] = __ddiv_rn(__dmul_rn(dr[k], r), dist);
            }
            s2 = __dadd_rn(__dmul_rn(dr[0], dr[0]), __dmul_rn(dr[1], dr[1]));
            if (DIM == 3) s2 = __dadd_rn(s2, __dmul_rn(dr[DIM - 1], dr[DIM - 1]));
            dist = sqrt(s2);
            const double ncost0 = __dadd_rn(cost[ni], dist);
#pragma unroll
            for (int k = 0; k < DIM; ++k) npos[k] = __dadd_rn(from[k], dr[k]);

            // ---- can_connect_nodes: warps 0..3 take one round of 25-32 samples each; the verdict is combined by the
            // barrier that publishes the near mask, which is computed speculatively (the test passes ~95 % of the time)
            const Segment<DIM> sg0 = segment_setup<DIM>(from, npos);
            bool my_blocked = sg0.same;
            if (warp < 4 && !sg0.same) {
                const int i = warp * kWarp + lane;
                my_blocked = i < 100 && sample_blocked<DIM, GS>(sg0, i, grid, shape);
            }

            // ---- find_nearby_nodes: euclid(new, node) <= r, as a bit mask in insertion order ----
            const double r2 = __dmul_rn(r, r);
            const double r2lo = __dmul_rn(r2, NRRT_K_LO), r2hi = __dmul_rn(r2, NRRT_K_HI);
            // integer new node against integer nodes: s is an exact integer and sqrt is monotone, so the reference's test
            // fl(sqrt(s)) <= r is s <= s_max with s_max = the largest integer that passes it (found once per iteration)
            bool q_exact;
            const PW qw = encode_pos<DIM>(npos, &q_exact);
            int qi[DIM];
#pragma unroll
            for (int k = 0; k < DIM; ++k) qi[k] = (int)(int16_t)(uint16_t)(((unsigned long long)qw >> (16 * k)) & 0xffffu);
            uint32_t s_max = 0;
            bool s_any = false;  // false: no integer passes (cannot happen for r >= 0, kept for safety)
            if (q_exact) {
                long long ci = (long long)fmin(floor(r2hi), 4294967295.0);
                for (; ci >= 0; --ci) {
                    const double cd = (double)ci;
                    if ((cd <= r2lo) ? true : ((cd >= r2hi) ? false : (sqrt(cd) <= r))) break;
                }
                s_any = ci >= 0;
                s_max = (uint32_t)(ci >= 0 ? ci : 0);
            }
            for (int base = 0; base < n; base += NT) {
                const int j = base + tid;
                bool in = false;
                if (j < n) {
                    const PW w = pw[j];
                    const int k0 = (int)(int16_t)(uint16_t)(w & 0xffffu);
                    if (q_exact && k0 != kPosSentinel) {
                        const int d0 = k0 - qi[0];
                        const int d1 = (int)(int16_t)(uint16_t)((w >> 16) & 0xffffu) - qi[1];
                        uint32_t sq = (uint32_t)(d0 * d0) + (uint32_t)(d1 * d1);
                        if (DIM == 3) {
                            const int d2 = (int)(int16_t)(uint16_t)(((unsigned long long)w >> 32) & 0xffffu) - qi[DIM - 1];
                            sq += (uint32_t)(d2 * d2);
                        }
                        in = s_any && sq <= s_max;
                    } else {
                        double pj[DIM];
                        load_pos<DIM>(pw, gpos, j, pj);
                        const double s = nrrt_sqdist<DIM>(npos, pj);
                        in = (s <= r2lo) ? true : ((s >= r2hi) ? false : (sqrt(s) <= r));
                    }
                }
                const uint32_t m = __ballot_sync(0xffffffffu, in);
                if (lane == 0 && (base >> 5) + warp < a.mask_words) mask[(base >> 5) + warp] = m;
            }
            const bool connect = __syncthreads_or(my_blocked ? 1 : 0) == 0;  // (2)
            if (TRACE && tid == 0) {
                if (trace_n < a.out.trace_cap) {
                    int32_t* t = trace + 4 * trace_n;
                    t[0] = iters; t[1] = 0; t[2] = ni; t[3] = connect ? 1 : 0;
                }
                ++trace_n;
            }
            if (!connect) continue;
            const int nwords = (n + 31) >> 5;

            // ---- loop A: choose parent.  Result = lexicographic min (cost, index) over free candidates < ncost0 ----
            double wbest = ncost0;
            int wpar = -1;
            // this warp's mask words: w = w_first + w_stride * t.  Lane t fetches word t of each block of 32 and the
            // non-empty ones are visited in order (most of the ~20 words per warp are empty).
            const int w_first = TRACE ? 0 : warp, w_stride = TRACE ? 1 : NW;
            const int my_words = (TRACE && warp != 0) ? 0 : (nwords - w_first + w_stride - 1) / w_stride;
            for (int tb = 0; tb < my_words; tb += kWarp) {
              const uint32_t lane_word = (tb + lane < my_words) ? mask[w_first + w_stride * (tb + lane)] : 0u;
              uint32_t nz = __ballot_sync(0xffffffffu, lane_word != 0u);
              while (nz) {
                const int t = __ffs(nz) - 1;
                nz &= nz - 1;
                const int w = w_first + w_stride * (tb + t);
                const uint32_t m = __shfl_sync(0xffffffffu, lane_word, t);
                const int j = (w << 5) + lane;
                const bool has = (m >> lane) & 1u;
                double pj[DIM], cj = INFINITY;
#pragma unroll
                for (int k = 0; k < DIM; ++k) pj[k] = 0.0;
                if (has) {
                    load_pos<DIM>(pw, gpos, j, pj);
                    cj = __dadd_rn(cost[j], sqrt(nrrt_sqdist<DIM>(pj, npos)));
                }
                uint32_t cand = __ballot_sync(0xffffffffu, has && cj < ncost0);
                while (cand) {
                    const int b = __ffs(cand) - 1;
                    cand &= cand - 1;
                    double p1[DIM];
#pragma unroll
                    for (int k = 0; k < DIM; ++k) p1[k] = __shfl_sync(0xffffffffu, pj[k], b);
                    const double cb = __shfl_sync(0xffffffffu, cj, b);
                    if (!(cb < wbest)) continue;  // reference would not call is_collision_free here (uniform per warp)
                    const bool fr = collision_free_warp<DIM, GS>(p1, npos, grid, shape, lane);
                    if (TRACE && lane == 0) {
                        if (trace_n < a.out.trace_cap) {
                            int32_t* t = trace + 4 * trace_n;
                            t[0] = iters; t[1] = 1; t[2] = (w << 5) + b; t[3] = fr ? 1 : 0;
                        }
                        ++trace_n;
                    }
                    if (fr) { wbest = cb; wpar = (w << 5) + b; }
                }
              }
            }
            if (lane == 0) { slot_a[warp].v = wbest; slot_a[warp].j = wpar; }
            __syncthreads();  // (3)
            double ncost = ncost0;
            int npar = ni;
            bool have = false;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const double c = slot_a[w].v;
                const int j = slot_a[w].j;
                if (j >= 0 && (!have || c < ncost || (c == ncost && j < npar))) { ncost = c; npar = j; have = true; }
            }

            // ---- loop B: rewire neighbours through the new node (index n); costs are not propagated ----
            for (int tb = 0; tb < my_words; tb += kWarp) {
              const uint32_t lane_word = (tb + lane < my_words) ? mask[w_first + w_stride * (tb + lane)] : 0u;
              uint32_t nz = __ballot_sync(0xffffffffu, lane_word != 0u);
              while (nz) {
                const int t = __ffs(nz) - 1;
                nz &= nz - 1;
                const int w = w_first + w_stride * (tb + t);
                const uint32_t m = __shfl_sync(0xffffffffu, lane_word, t);
                const int j = (w << 5) + lane;
                const bool has = (m >> lane) & 1u;
                double pj[DIM], cj = 0.0, cur = 0.0;
#pragma unroll
                for (int k = 0; k < DIM; ++k) pj[k] = 0.0;
                if (has) {
                    load_pos<DIM>(pw, gpos, j, pj);
                    cj = __dadd_rn(ncost, sqrt(nrrt_sqdist<DIM>(npos, pj)));
                    cur = cost[j];
                }
                uint32_t cand = __ballot_sync(0xffffffffu, has && cj < cur);
                while (cand) {
                    const int b = __ffs(cand) - 1;
                    cand &= cand - 1;
                    double p2[DIM];
#pragma unroll
                    for (int k = 0; k < DIM; ++k) p2[k] = __shfl_sync(0xffffffffu, pj[k], b);
                    const bool fr = collision_free_warp<DIM, GS>(npos, p2, grid, shape, lane);
                    if (TRACE && lane == 0) {
                        if (trace_n < a.out.trace_cap) {
                            int32_t* t = trace + 4 * trace_n;
                            t[0] = iters; t[1] = 2; t[2] = (w << 5) + b; t[3] = fr ? 1 : 0;
                        }
                        ++trace_n;
                    }
                    if (fr && lane == b) { parent[j] = n; cost[j] = cj; }
                }
              }
            }

            // ---- goal_reached + append (rrt_star.py:156-161, :201-209) ----
            const double dg = sqrt(nrrt_sqdist<DIM>(npos, goal));
            if (dg < best) { best = dg; best_node = n; }
            const bool reached = dg <= thr;
            if (tid == 0) {
                double fin[DIM];
#pragma unroll
                for (int k = 0; k < DIM; ++k) fin[k] = reached ? goal[k] : npos[k];  // goal_node.position = self.goal
                bool exact;
                pw[n] = encode_pos<DIM>(fin, &exact);
                if (!exact) {
#pragma unroll
                    for (int k = 0; k < DIM; ++k) gpos[(size_t)n * DIM + k] = fin[k];
                }
                cost[n] = ncost; parent[n] = npar;
            }
            if (reached) { goal_node = n; break; }
            // no barrier here: the next scan takes the new node from registers, and every later reader of pos[n],
            // cost[n], parent[n], mask or slot_a sits behind barrier (1) of the next iteration
#pragma unroll
            for (int k = 0; k < DIM; ++k) pend[k] = npos[k];
            pend_w = qw;  // (npos is not the goal here: a reached goal leaves the loop above)
            pend_j = n;
            ++n;
        }
        __syncthreads();

        // ---- epilogue: scalars, find_path + calculate_length, optional tree dump ----
        const int n_tot = n + (goal_node >= 0 ? 1 : 0);
        for (int j = tid; j < n_tot; j += NT) {  // tree dump: the integer nodes join the fp64 ones already in `pos`
            const PW w = pw[j];
            if ((int)(int16_t)(uint16_t)(w & 0xffffu) != kPosSentinel) {
                double pj[DIM];
                load_pos<DIM>(pw, gpos, j, pj);
#pragma unroll
                for (int k = 0; k < DIM; ++k) gpos[(size_t)j * DIM + k] = pj[k];
            }
        }
        if (tid == 0) {
            a.out.n_nodes[task] = n;
            a.out.iterations[task] = iters;
            a.out.status[task] = goal_node >= 0 ? NRRT_SOLVED
                                                : (best_node >= 0 ? NRRT_MAX_ITER_BEST_EFFORT : NRRT_NO_NODE_CONNECTED);
            a.out.goal_node[task] = goal_node;
            a.out.best_node[task] = best_node;
            a.out.best_dist[task] = best;
            if (TRACE && a.out.trace_n) a.out.trace_n[task] = trace_n;
            if (a.out.path_n) {
                const int end = goal_node >= 0 ? goal_node : best_node;
                int m = 0;
                for (int j = end; j >= 0 && m <= n_tot; j = parent[j]) ++m;  // find_path: walk to the root
                if (m > n_tot) m = 0;  // parent cycle (the reference would never return): report no path
                a.out.path_n[task] = m;
                double len = 0.0;
                if (a.out.path_idx && m <= a.p.path_cap) {
                    int32_t* pi = a.out.path_idx + (size_t)task * a.p.path_cap;
                    int q = m;
                    for (int j = end; j >= 0; j = parent[j]) pi[--q] = j;  // path.reverse()
                    for (int q2 = 0; q2 + 1 < m; ++q2) {                   // calculate_length: forward fp64 sum
                        double u[DIM], v[DIM];
                        load_pos<DIM>(pw, gpos, pi[q2], u);
                        load_pos<DIM>(pw, gpos, pi[q2 + 1], v);
                        len = __dadd_rn(len, sqrt(nrrt_sqdist<DIM>(u, v)));
                    }
                } else if (m > 0) {
                    len = NAN;  // path longer than path_cap: length not reported
                }
                if (a.out.path_len) a.out.path_len[task] = len;
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Stand-alone batched primitives behind the reference's individually callable methods (same device functions as the
// planner kernel): is_collision_free, find_nearest_neighbor, find_nearby_nodes.
// ------------------------------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(128) collision_kernel(const uint32_t* __restrict__ occ_bits, int words_per_map,
                                                        const int32_t* __restrict__ seg_map, const double* __restrict__ p1,
                                                        const double* __restrict__ p2, int S, int s0, int s1, int s2,
                                                        int32_t* __restrict__ verdict) {
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= S) return;
    const int shape[3] = {s0, s1, s2};
    double a[DIM], b[DIM];
#pragma unroll
    for (int k = 0; k < DIM; ++k) { a[k] = p1[(size_t)s * DIM + k]; b[k] = p2[(size_t)s * DIM + k]; }
    const bool fr = collision_free_warp<DIM>(a, b, occ_bits + (size_t)seg_map[s] * words_per_map, shape, lane);
    if (lane == 0) verdict[s] = fr ? 1 : 0;
}

// one CTA per query: nearest node index (strict <, lowest index on ties of the sqrt'ed distance) and the near mask
template <int DIM>
__global__ void __launch_bounds__(256) search_kernel(const double* __restrict__ nodes, int n, const double* __restrict__ queries,
                                                     const double* __restrict__ radii, int32_t* __restrict__ nearest,
                                                     uint32_t* __restrict__ near_mask, int mask_words) {
    __shared__ Slot slots[8];
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double qp[DIM];
#pragma unroll
    for (int k = 0; k < DIM; ++k) qp[k] = queries[(size_t)q * DIM + k];
    double bs = INFINITY;
    int bj = 0x7fffffff;
    for (int j = tid; j < n; j += 256) {
        double pj[DIM];
#pragma unroll
        for (int k = 0; k < DIM; ++k) pj[k] = nodes[(size_t)j * DIM + k];
        const double s = nrrt_sqdist<DIM>(pj, qp);
        if (s < bs) {
            bool take = true;
            if (s > __dmul_rn(bs, NRRT_K_LO)) take = nn_take_slow(s, bs);
            if (take) { bs = s; bj = j; }
        }
    }
    nn_warp_reduce(bs, bj);
    if (lane == 0) { slots[warp].v = bs; slots[warp].j = bj; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; ++w)
            if (!nn_a_wins(bs, bj, slots[w].v, slots[w].j)) { bs = slots[w].v; bj = slots[w].j; }
        nearest[q] = n > 0 ? bj : -1;
    }
    if (near_mask) {
        const double r = radii[q];
        const double r2 = __dmul_rn(r, r);
        const double r2lo = __dmul_rn(r2, NRRT_K_LO), r2hi = __dmul_rn(r2, NRRT_K_HI);
        for (int base = 0; base < n; base += 256) {
            const int j = base + tid;
            bool in = false;
            if (j < n) {
                double pj[DIM];
#pragma unroll
                for (int k = 0; k < DIM; ++k) pj[k] = nodes[(size_t)j * DIM + k];
                const double s = nrrt_sqdist<DIM>(qp, pj);
                in = (s <= r2lo) ? true : ((s >= r2hi) ? false : (sqrt(s) <= r));
            }
            const uint32_t m = __ballot_sync(0xffffffffu, in);
            if (lane == 0 && (base >> 5) + warp < mask_words) near_mask[(size_t)q * mask_words + (base >> 5) + warp] = m;
        }
    }
}

}  // namespace

extern "C" int nrrt_collision_free(int dim, const int32_t* shape, const uint32_t* occ_bits, const int32_t* seg_map,
                                   const double* p1, const double* p2, int S, int32_t* verdict, void* stream) {
    NRRT_REQUIRE((dim == 2 || dim == 3) && shape && occ_bits && seg_map && p1 && p2 && verdict, "bad collision arguments");
    if (S <= 0) return NRRT_OK;
    const int64_t cells = (int64_t)shape[0] * shape[1] * (dim == 3 ? shape[2] : 1);
    const int wpm = (int)nrrt_words_per_map(cells);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dim == 2) collision_kernel<2><<<(S + 3) / 4, 128, 0, st>>>(occ_bits, wpm, seg_map, p1, p2, S, shape[0], shape[1], 1, verdict);
    else collision_kernel<3><<<(S + 3) / 4, 128, 0, st>>>(occ_bits, wpm, seg_map, p1, p2, S, shape[0], shape[1], shape[2], verdict);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_tree_search(int dim, const double* nodes, int n, const double* queries, const double* radii, int Q,
                                int32_t* nearest, uint32_t* near_mask, void* stream) {
    NRRT_REQUIRE((dim == 2 || dim == 3) && nodes && queries && nearest && n >= 0, "bad search arguments");
    NRRT_REQUIRE(near_mask == nullptr || radii != nullptr, "near_mask needs radii");
    if (Q <= 0) return NRRT_OK;
    const int mask_words = (n + 31) / 32;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dim == 2) search_kernel<2><<<Q, 256, 0, st>>>(nodes, n, queries, radii, nearest, near_mask, mask_words);
    else search_kernel<3><<<Q, 256, 0, st>>>(nodes, n, queries, radii, nearest, near_mask, mask_words);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

namespace {

struct SmemPlan {
    size_t bytes;
    int cap, mask_words, nt;
};

SmemPlan smem_plan(const NrrtPlanParams& p, int words_per_map, bool grid_in_smem = true) {
    SmemPlan s;
    s.nt = p.dim == 2 ? 128 : 256;
    s.cap = (p.max_iterations + 2 + 3) & ~3;  // packed positions end on a 16-byte boundary (the grid is staged with 16-byte stores)
    s.mask_words = (s.cap + 31) / 32;
    const int nw = s.nt / 32;
    s.bytes = (size_t)s.cap * (p.dim == 2 ? 4 : 8) + (grid_in_smem ? (size_t)((words_per_map + 3) & ~3) * 4 : 0) +
              (size_t)((s.mask_words + 3) & ~3) * 4 + (size_t)3 * nw * sizeof(Slot);
    return s;
}

template <int DIM, int NT, bool TRACE, bool GS = true, int MINB = (DIM == 2 ? 4 : 2)>
int launch_plan(const PlanArgs& a, size_t smem, cudaStream_t st) {
    auto kern = plan_kernel<DIM, NT, TRACE, GS, MINB>;
    NRRT_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0, per_sm = 0;
    NRRT_CUDA_TRY(cudaGetDevice(&dev));
    NRRT_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    NRRT_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem));
    if (per_sm < 1) return nrrt_fail(NRRT_ERR_UNSUPPORTED, "planner CTA needs %zu B of shared memory: does not fit", smem);
    const int grid = a.B < sms * per_sm ? a.B : sms * per_sm;  // persistent CTAs pulling tasks from the queue
    kern<<<grid, NT, smem, st>>>(a);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

}  // namespace

extern "C" int nrrt_plan_smem_bytes(const NrrtPlanParams* p, size_t* bytes, int* ctas_per_sm) {
    NRRT_REQUIRE(p && (p->dim == 2 || p->dim == 3), "dim must be 2 or 3");
    const int64_t cells = (int64_t)p->shape[0] * p->shape[1] * (p->dim == 3 ? p->shape[2] : 1);
    const SmemPlan s = smem_plan(*p, (int)nrrt_words_per_map(cells), false);  // the production kernels (map bits through L1)
    if (bytes) *bytes = s.bytes;
    if (ctas_per_sm) {
        const int by_smem = (int)(233472 / (s.bytes + 1024)), by_regs = p->dim == 2 ? 8 : 4;
        *ctas_per_sm = by_smem < by_regs ? by_smem : by_regs;
    }
    return NRRT_OK;
}

extern "C" int nrrt_plan_batch(const NrrtPlanParams* p, const uint32_t* occ_bits, const int32_t* task_to_map,
                               const double* radius, const int16_t* samples, const int32_t* starts, const int32_t* goals,
                               const int32_t* draw_status, const NrrtPlanOut* out, int B, int32_t* queue_counter,
                               void* stream) {
    NRRT_REQUIRE(p && out, "null params/out");
    NRRT_REQUIRE(p->dim == 2 || p->dim == 3, "dim must be 2 or 3, got %d", p->dim);
    NRRT_REQUIRE(B >= 0, "negative batch");
    NRRT_REQUIRE(p->max_iterations >= 1 && p->max_iterations <= 32000, "max_iterations out of range: %d", p->max_iterations);
    NRRT_REQUIRE(p->node_stride >= p->max_iterations + 2, "node_stride %d < max_iterations + 2", p->node_stride);
    if (B == 0) return NRRT_OK;
    NRRT_REQUIRE(occ_bits && task_to_map && radius && samples && starts && goals && queue_counter, "null input pointer");
    NRRT_REQUIRE(out->pos && out->cost && out->parent && out->n_nodes && out->iterations && out->status && out->goal_node &&
                     out->best_node && out->best_dist, "null required output pointer");
    for (int k = 0; k < p->dim; ++k) NRRT_REQUIRE(p->shape[k] >= 1 && p->shape[k] <= 32767, "bad shape[%d]=%d", k, p->shape[k]);
    const int64_t cells = (int64_t)p->shape[0] * p->shape[1] * (p->dim == 3 ? p->shape[2] : 1);
    NRRT_REQUIRE(cells < (int64_t)1 << 31, "map too large");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

    PlanArgs a;
    a.p = *p; a.occ_bits = occ_bits; a.task_to_map = task_to_map; a.radius = radius; a.samples = samples;
    a.starts = starts; a.goals = goals; a.draw_status = draw_status; a.out = *out; a.B = B; a.queue = queue_counter;
    a.words_per_map = (int)nrrt_words_per_map(cells);
    const SmemPlan s = smem_plan(*p, a.words_per_map);
    a.cap = s.cap; a.mask_words = s.mask_words;
    const bool tracing = out->trace != nullptr && out->trace_cap > 0;
    const size_t need = tracing ? s.bytes : smem_plan(*p, a.words_per_map, false).bytes;
    if (need > 232448) return nrrt_fail(NRRT_ERR_UNSUPPORTED, "planner needs %zu B shared memory (> 227 KiB): map or max_iterations too large", need);
    NRRT_CUDA_TRY(cudaMemsetAsync(queue_counter, 0, sizeof(int32_t), st));
    // Production kernels: map bits through L1, 8 (2D) / 4 (3D) CTAs per SM (see the layout note at the top of the file).
    // The TRACE kernels (parity tests only) keep the staged copy and the old occupancy: both paths are compared with the
    // reference's recorded trees, so the two grid paths check each other.
    if (!tracing) {
        const SmemPlan g = smem_plan(*p, a.words_per_map, false);
        if (p->dim == 2) return launch_plan<2, 128, false, false, 8>(a, g.bytes, st);
        return launch_plan<3, 256, false, false, 4>(a, g.bytes, st);
    }
    if (p->dim == 2) return launch_plan<2, 128, true>(a, s.bytes, st);
    return launch_plan<3, 256, true>(a, s.bytes, st);
}
