// gridpath.cu — batched exact shortest grid paths: the "A*" column of the reference's results table and its task filter.
//
// Replaces generate_paths.astar (generate_paths.py:11-74) and ThreeD_generate_paths.astar (ThreeD_generate_paths.py:14-97)
// for thousands of (map, start, goal) tasks at once.  2D: 8-connected moves, cost 1 / sqrt(2) (:66-72), a diagonal step
// only when BOTH orthogonal neighbours are free (:45-59).  3D: 26-connected moves, cost 1 / sqrt(2) / sqrt(3) (:83-97); the
// reference compares the move OFFSETS with 255 (`dx != 255`, :55-72), which is always true, so EVERY move -- axis moves
// included -- needs its three axis-projected cells free (a projection with a zero offset is the current cell itself).
// No path -> the task is deleted (generate_paths.py:111-113).  A* with the (consistent) Euclidean heuristic returns a
// shortest path, and every shortest path has the same numbers of steps of each kind (1, sqrt(2), sqrt(3) are rationally
// independent), so the length the reference reports (calculate_length, test_network_and_algorithms.py:52-56) is
// n1 + n2 * sqrt(2) + n3 * sqrt(3) up to the order of its additions.  The kernel computes exactly those integers.
//
// Algorithm: label-setting Dijkstra by unit-width buckets (Dial).  Every edge is >= 1 = the bucket width, so no relaxation
// out of bucket k lands in bucket k: all nodes of a bucket are final when it is processed and can be expanded in parallel;
// relaxations land in buckets k+1 and k+2 only (max edge < 2), so three shared-memory queues in rotation suffice.
// One task per CTA (persistent CTAs, atomic task queue).  Node labels live in a per-CTA slot of global memory (8 B per
// cell, 2 MB at 512^2 / 4 MB at 80^3: L2-resident) as one 64-bit word [fixed-point cost | diagonal step counts] (see Fx
// below) updated with atomicMin: the cost orders candidates and carries the step counts with it.  The occupancy grid is
// staged in shared memory (bit-packed, as in the planner).  A queue overflow degrades that bucket to a scan of all cells.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>

#include "nrrt_common.cuh"

namespace {

// 2D (512^2: 32 KB grid): 256 threads and 6144-entry queues = 104 KB, two CTAs per SM -- the kernel is barrier / latency
// bound (ncu: 62 % of stall samples at the two block barriers of a bucket step, issue slots 7 % busy), so tasks in flight
// per SM is what buys throughput.  3D (80^3: 64 KB grid, wide fronts): 512 threads, 12288-entry queues, one CTA per SM.
template <int DIM> struct GP { static constexpr int NT = DIM == 2 ? 256 : 512, QCAP = DIM == 2 ? 6144 : 12288, PER_SM = DIM == 2 ? 2 : 1; };
constexpr unsigned long long kInf = ~0ull;
// Label = [fixed-point cost | step counts of the diagonal kinds]; the number of unit steps is recovered from the cost
// (n1 = (cost - n2 K2 - n3 K3) >> F, exact), which leaves the cost 38 (2D) / 31 (3D) fraction bits: the fixed-point cost
// differs from the true one by < n * 2^-(F+1), i.e. < 1.5e-8 (2D, n <= 8191) / < 7e-7 (3D) at the limits below.  2D: two
// different step-count pairs differ by |da + db sqrt(2)| >= 1 / (|da| + |db| sqrt(2)) > 4e-5 (norm form a^2 - 2 b^2 != 0),
// so the fixed-point order IS the exact order.  3D: not provable from the norm bound (1e-12 at the limit); for paths of a
// few hundred steps (80^3 maps) the smallest gap of candidate costs is ~1e-6 against an error < 1e-7.
// Limits: cost < 8191 (2D) / 1771 (3D: n2 <= 2047, n3 <= 1023 fit their fields); a search that reaches the limit reports
// n_steps = -2 and length = +inf ("limit") -- never "no path".
template <int DIM> struct Fx {
    static constexpr int F = DIM == 2 ? 38 : 31, PAY = DIM == 2 ? 13 : 21, LIMIT = DIM == 2 ? 8191 : 1771;
    static constexpr unsigned long long K1 = 1ull << F;
    static constexpr unsigned long long K2 = DIM == 2 ? 388736063997ull : 3037000500ull;  // round(sqrt(2) * 2^F)
    static constexpr unsigned long long K3 = 3719550787ull;                             // round(sqrt(3) * 2^31)
};

struct GridPathArgs {
    const uint32_t* occ_bits;
    int words_per_map;
    const int32_t* task_to_map;
    const int32_t* starts;  // [B][DIM]: (y, x) in 2D (generate_paths.py:81-87), (x, y, z) = array axes in 3D
    const int32_t* goals;
    int B, S0, S1, S2;         // grid shape = array axes (S2 = 1 in 2D)
    unsigned long long* work;  // [gridDim.x][cells]
    int32_t* queue;            // zeroed task counter
    int axis_only;             // 1: axis moves only (4- / 6-connected reachability); 0: the reference's A* moves
    int32_t* n_steps;          // [B][3]: steps of cost 1, sqrt(2), sqrt(3); -1 = no path
    double* length;            // [B], NaN = no path
    int32_t* paths;            // optional [B][path_cap][DIM]: one shortest path, start -> goal, array-axis order
    int32_t* path_n;           // optional [B]: points on it (steps + 1), 0 = no path, -(steps + 1) = longer than path_cap
    int path_cap;
};

template <int DIM>
__device__ __forceinline__ unsigned long long make_key(unsigned a, unsigned b, unsigned c) {
    const unsigned long long cost = a * Fx<DIM>::K1 + b * Fx<DIM>::K2 + (DIM == 3 ? c * Fx<DIM>::K3 : 0ull);
    const unsigned long long pay = DIM == 2 ? (unsigned long long)b : (((unsigned long long)b << 10) | c);
    return (cost << Fx<DIM>::PAY) | pay;
}
template <int DIM>
__device__ __forceinline__ int key_bucket(unsigned long long key) { return (int)(key >> (Fx<DIM>::PAY + Fx<DIM>::F)); }
template <int DIM>
__device__ __forceinline__ void key_counts(unsigned long long key, unsigned& a, unsigned& b, unsigned& c) {
    if (DIM == 2) { b = (unsigned)key & 0x1fffu; c = 0; }
    else { b = (unsigned)(key >> 10) & 0x7ffu; c = (unsigned)key & 0x3ffu; }
    a = (unsigned)(((key >> Fx<DIM>::PAY) - b * Fx<DIM>::K2 - (DIM == 3 ? c * Fx<DIM>::K3 : 0ull)) >> Fx<DIM>::F);
}

template <int DIM>
__global__ void __launch_bounds__(GP<DIM>::NT, GP<DIM>::PER_SM) grid_path_kernel(const GridPathArgs a) {
    constexpr int kNT = GP<DIM>::NT, kQCap = GP<DIM>::QCAP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* grid = reinterpret_cast<uint32_t*>(smem_raw);                 // [words_per_map] (padded to 4)
    uint32_t* q = grid + ((a.words_per_map + 3) & ~3);                      // [3][kQCap]
    __shared__ int cnt[3], ovf[3], s_task, s_state;
    const int tid = threadIdx.x;
    const int S12 = a.S1 * a.S2;
    const int cells = a.S0 * S12;
    unsigned long long* dist = a.work + (size_t)blockIdx.x * cells;
    auto is_free = [&](int i) -> bool { return (grid[i >> 5] >> (i & 31)) & 1u; };
    int cur_map = -1;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_task = atomicAdd(a.queue, 1);
        __syncthreads();
        const int task = s_task;
        if (task >= a.B) break;
        const int map = a.task_to_map[task];
        if (map != cur_map) {
            for (int i = tid; i < a.words_per_map; i += kNT) grid[i] = __ldg(a.occ_bits + (size_t)map * a.words_per_map + i);
            cur_map = map;
        }
        for (int i = tid; i < cells; i += kNT) dist[i] = kInf;
        int s = 0, g = 0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
            const int stride = d == 0 ? S12 : (d == 1 ? a.S2 : 1);
            s += a.starts[task * DIM + d] * (DIM == 2 ? (d == 0 ? a.S1 : 1) : stride);
            g += a.goals[task * DIM + d] * (DIM == 2 ? (d == 0 ? a.S1 : 1) : stride);
        }
        if (tid < 3) { cnt[tid] = 0; ovf[tid] = 0; }
        __syncthreads();
        if (tid == 0) { dist[s] = make_key<DIM>(0, 0, 0); q[0] = (uint32_t)s; cnt[0] = 1; s_state = 0; }
        __syncthreads();
        for (int k = 0; k < Fx<DIM>::LIMIT; ++k) {
            const int cur = k % 3;
            const bool scan = ovf[cur] != 0;
            const int n = scan ? cells : min(cnt[cur], kQCap);
            for (int i = tid; i < n; i += kNT) {
                const int u = scan ? i : (int)q[cur * kQCap + i];
                const unsigned long long key = __ldcg(dist + u);  // labels are updated by L2 atomics: never read them through L1
                if (key == kInf || key_bucket<DIM>(key) != k) continue;  // stale entry: the node moved to an earlier bucket
                unsigned n1, n2, n3;
                key_counts<DIM>(key, n1, n2, n3);
                // coordinates along the array axes; 2D: (c0, c1) = (y, x) on an (S0, S1) grid, c2 = 0
                const int c0 = u / S12, r = u - c0 * S12;
                const int c1 = DIM == 2 ? r : r / a.S2, c2 = DIM == 2 ? 0 : r - c1 * a.S2;
                const int st1 = DIM == 2 ? 1 : a.S2;
#pragma unroll
                for (int d0 = -1; d0 <= 1; ++d0)
#pragma unroll
                    for (int d1 = -1; d1 <= 1; ++d1)
#pragma unroll
                        for (int d2 = (DIM == 3 ? -1 : 0); d2 <= (DIM == 3 ? 1 : 0); ++d2) {
                            const int nz = (d0 != 0) + (d1 != 0) + (d2 != 0);
                            if (nz == 0 || (a.axis_only && nz != 1)) continue;
                            const int e0 = c0 + d0, e1 = c1 + d1, e2 = c2 + d2;
                            if (e0 < 0 || e0 >= a.S0 || e1 < 0 || e1 >= a.S1 || e2 < 0 || e2 >= a.S2) continue;
                            const int v = e0 * S12 + e1 * st1 + e2;
                            if (!is_free(v)) continue;
                            if (DIM == 2) {  // no corner cutting: both orthogonal neighbours of a diagonal step are free
                                if (nz == 2 && (!is_free(e0 * S12 + c1) || !is_free(c0 * S12 + e1))) continue;
                            } else {         // all three axis-projected cells are free (zero offset: the current cell)
                                if (!is_free(e0 * S12 + c1 * st1 + c2) || !is_free(c0 * S12 + e1 * st1 + c2) ||
                                    !is_free(c0 * S12 + c1 * st1 + e2))
                                    continue;
                            }
                            const unsigned long long nk = make_key<DIM>(n1 + (nz == 1), n2 + (nz == 2), n3 + (nz == 3));
                            if (nk < __ldcg(dist + v)) {  // cheap pre-test: most neighbours are already labelled better
                                const unsigned long long old = atomicMin(dist + v, nk);
                                if (nk < old) {
                                    const int bkt = key_bucket<DIM>(nk) % 3;
                                    const int pos = atomicAdd(&cnt[bkt], 1);
                                    if (pos < kQCap) q[bkt * kQCap + pos] = (uint32_t)v;
                                    else ovf[bkt] = 1;
                                }
                            }
                        }
            }
            __syncthreads();
            if (tid == 0) {
                const unsigned long long gk = __ldcg(dist + g);
                const int q1 = (k + 1) % 3, q2 = (k + 2) % 3;
                cnt[cur] = 0; ovf[cur] = 0;
                if (gk != kInf && key_bucket<DIM>(gk) <= k) s_state = 1;                                 // goal label final
                else if (cnt[q1] == 0 && cnt[q2] == 0 && !ovf[q1] && !ovf[q2]) s_state = 2;          // nothing left: no path
            }
            __syncthreads();
            if (s_state) break;
        }
        // one shortest path by backtracking the final labels (warp 0; lane = neighbour): the predecessor of a cell with step
        // counts (n1, n2, n3) is a neighbour whose counts are one step of the move's kind smaller and from which the move is
        // legal.  Labels at or below the goal's are final.  Among equal-cost paths the reference's A* returns the one its
        // heap order finds (generate_paths.py:11-74); this one takes the lowest neighbour index -- same length, same kinds.
        if (a.paths && tid < 32 && s_state == 1) {
            const unsigned long long gk = __ldcg(dist + g);
            unsigned n1, n2, n3;
            key_counts<DIM>(gk, n1, n2, n3);
            const int steps = (int)(n1 + n2 + n3);
            int32_t* out = a.paths + (size_t)task * a.path_cap * DIM;
            if (steps + 1 > a.path_cap) {
                if (tid == 0) a.path_n[task] = -(steps + 1);
            } else {
                if (tid == 0) a.path_n[task] = steps + 1;
                int u = g;
                const int st1 = DIM == 2 ? 1 : a.S2;
                for (int i = steps; i >= 0; --i) {
                    const int c0 = u / S12, r = u - c0 * S12;
                    const int c1 = DIM == 2 ? r : r / a.S2, c2 = DIM == 2 ? 0 : r - c1 * a.S2;
                    if (tid == 0) {
                        out[i * DIM] = c0; out[i * DIM + 1] = c1;
                        if (DIM == 3) out[i * DIM + 2] = c2;
                    }
                    if (i == 0) break;
                    key_counts<DIM>(__ldcg(dist + u), n1, n2, n3);
                    bool ok = false;
                    int v = 0;
                    if (tid < (DIM == 2 ? 9 : 27)) {
                        const int d0 = tid / (DIM == 2 ? 3 : 9) - 1, d1 = (tid / (DIM == 2 ? 1 : 3)) % 3 - 1, d2 = DIM == 2 ? 0 : tid % 3 - 1;
                        const int nz = (d0 != 0) + (d1 != 0) + (d2 != 0);
                        const int e0 = c0 - d0, e1 = c1 - d1, e2 = c2 - d2;          // candidate predecessor: it moved by (d0, d1, d2)
                        if (nz != 0 && !(a.axis_only && nz != 1) && e0 >= 0 && e0 < a.S0 && e1 >= 0 && e1 < a.S1 && e2 >= 0 && e2 < a.S2) {
                            v = e0 * S12 + e1 * st1 + e2;
                            bool legal = is_free(v);
                            if (DIM == 2) legal = legal && (nz != 2 || (is_free(c0 * S12 + e1) && is_free(e0 * S12 + c1)));
                            else legal = legal && is_free(c0 * S12 + e1 * st1 + e2) && is_free(e0 * S12 + c1 * st1 + e2) &&
                                         is_free(e0 * S12 + e1 * st1 + c2);       // the move's projected cells, seen from the predecessor
                            if (legal) {
                                const unsigned long long vk = __ldcg(dist + v);
                                if (vk != kInf) {
                                    unsigned m1, m2, m3;
                                    key_counts<DIM>(vk, m1, m2, m3);
                                    ok = m1 + (nz == 1) == n1 && m2 + (nz == 2) == n2 && m3 + (nz == 3) == n3;
                                }
                            }
                        }
                    }
                    const unsigned vote = __ballot_sync(0xffffffffu, ok);
                    if (vote == 0) {             // cannot happen for a final label; never loop on a broken chain
                        if (tid == 0) a.path_n[task] = 0;
                        break;
                    }
                    u = __shfl_sync(0xffffffffu, v, __ffs(vote) - 1);
                }
            }
        } else if (a.paths && tid == 0 && s_state != 1) {
            a.path_n[task] = 0;
        }
        if (tid == 0) {
            const unsigned long long gk = __ldcg(dist + g);
            if (s_state == 1 && gk != kInf) {
                unsigned n1, n2, n3;
                key_counts<DIM>(gk, n1, n2, n3);
                a.n_steps[task * 3] = (int)n1; a.n_steps[task * 3 + 1] = (int)n2; a.n_steps[task * 3 + 2] = (int)n3;
                a.length[task] = (double)n1 + (double)n2 * 1.4142135623730951 + (double)n3 * 1.7320508075688772;
            } else {  // s_state 2: every reachable cell labelled, the goal is not among them; 0: the cost limit was reached
                a.n_steps[task * 3] = a.n_steps[task * 3 + 1] = a.n_steps[task * 3 + 2] = s_state == 2 ? -1 : -2;
                a.length[task] = s_state == 2 ? nan("") : __longlong_as_double(0x7ff0000000000000ll);
            }
        }
    }
}

}  // namespace

extern "C" int nrrt_grid_paths_ex(int dims, const uint32_t* occ_bits, int words_per_map, const int32_t* task_to_map,
                                  const int32_t* starts, const int32_t* goals, int B, int s0, int s1, int s2, int axis_only,
                                  void* work, int n_slots, int32_t* queue, int32_t* n_steps, double* length, int32_t* paths,
                                  int path_cap, int32_t* path_n, void* stream);

extern "C" int nrrt_grid_paths(int dims, const uint32_t* occ_bits, int words_per_map, const int32_t* task_to_map,
                               const int32_t* starts, const int32_t* goals, int B, int s0, int s1, int s2, int axis_only,
                               void* work, int n_slots, int32_t* queue, int32_t* n_steps, double* length, void* stream) {
    return nrrt_grid_paths_ex(dims, occ_bits, words_per_map, task_to_map, starts, goals, B, s0, s1, s2, axis_only, work, n_slots, queue,
                              n_steps, length, nullptr, 0, nullptr, stream);
}

extern "C" int nrrt_grid_paths_ex(int dims, const uint32_t* occ_bits, int words_per_map, const int32_t* task_to_map,
                                  const int32_t* starts, const int32_t* goals, int B, int s0, int s1, int s2, int axis_only,
                                  void* work, int n_slots, int32_t* queue, int32_t* n_steps, double* length, int32_t* paths,
                                  int path_cap, int32_t* path_n, void* stream) {
    NRRT_REQUIRE((paths == nullptr) == (path_n == nullptr) && (paths == nullptr || path_cap >= 1), "paths, path_n and path_cap go together");
    NRRT_REQUIRE(occ_bits && task_to_map && starts && goals && work && queue && n_steps && length, "null pointer");
    NRRT_REQUIRE(dims == 2 || dims == 3, "dims must be 2 or 3");
    if (dims == 2) s2 = 1;
    NRRT_REQUIRE(B >= 0 && s0 >= 1 && s1 >= 1 && s2 >= 1 && s0 <= 2048 && s1 <= 2048 && s2 <= 2048 &&
                     (long long)words_per_map * 32 >= (long long)s0 * s1 * s2 && (long long)s0 * s1 * s2 < (1ll << 30),
                 "bad grid shape");
    NRRT_REQUIRE(n_slots >= 1, "need at least one work slot");
    if (B == 0) return NRRT_OK;
    const size_t smem = (size_t)((words_per_map + 3) & ~3) * 4 + (size_t)3 * (dims == 2 ? GP<2>::QCAP : GP<3>::QCAP) * 4;
    NRRT_REQUIRE(smem <= 227 * 1024, "grid too large for shared memory (%zu bytes)", smem);
    GridPathArgs a;
    a.occ_bits = occ_bits; a.words_per_map = words_per_map; a.task_to_map = task_to_map; a.starts = starts; a.goals = goals;
    a.B = B; a.S0 = s0; a.S1 = s1; a.S2 = s2; a.axis_only = axis_only ? 1 : 0; a.work = (unsigned long long*)work; a.queue = queue;
    a.n_steps = n_steps; a.length = length;
    a.paths = paths; a.path_n = path_n; a.path_cap = path_cap;
    const int grid = n_slots < B ? n_slots : B;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dims == 2) {
        NRRT_CUDA_TRY(cudaFuncSetAttribute(grid_path_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        grid_path_kernel<2><<<grid, GP<2>::NT, smem, st>>>(a);
    } else {
        NRRT_CUDA_TRY(cudaFuncSetAttribute(grid_path_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        grid_path_kernel<3><<<grid, GP<3>::NT, smem, st>>>(a);
    }
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}
