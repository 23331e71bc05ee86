// gridpath.cu — batched exact shortest grid paths: the "A*" column of the reference's results table and its task filter.
//
// Replaces generate_paths.astar (generate_paths.py:11-74) for thousands of (map, start, goal) tasks at once: 8-connected
// moves on the 2D occupancy grid, cost 1 / sqrt(2) (:66-72), a diagonal step only when BOTH orthogonal neighbours are free
// (:45-59), no path -> the task is deleted (:111-113).  A* with the (consistent) Euclidean heuristic returns a shortest
// path, and every shortest path has the same numbers of straight and diagonal steps (1 and sqrt(2) are rationally
// independent), so the length the reference reports (calculate_length, test_network_and_algorithms.py:52-56) is
// n_straight + n_diagonal * sqrt(2) up to the order of its additions.  The kernel computes exactly those two integers.
//
// Algorithm: label-setting Dijkstra by unit-width buckets (Dial).  Every edge is >= 1 = the bucket width, so no relaxation
// out of bucket k lands in bucket k: all nodes of a bucket are final when it is processed and can be expanded in parallel;
// relaxations land in buckets k+1 and k+2 only (max edge < 2), so three shared-memory queues in rotation suffice.
// One task per CTA (persistent CTAs, atomic task queue).  Node labels live in a per-CTA slot of global memory (8 B per
// cell, 2 MB at 512^2: L2-resident) as one 64-bit word  [cost in 12.20 fixed point | n_straight:16 | n_diagonal:16]
// updated with atomicMin: the fixed-point cost a * 2^20 + b * round(sqrt(2) * 2^20) is an exact integer function of (a, b),
// orders candidates, and carries its step counts with it.  The occupancy grid is staged in shared memory (bit-packed, as
// in the planner).  A queue overflow (never seen at 512^2 with 12288 entries) degrades that bucket to a scan of all cells.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>

#include "nrrt_common.cuh"

namespace {

constexpr int kNT = 512;
constexpr int kQCap = 12288;
constexpr unsigned long long kInf = ~0ull;
constexpr unsigned long long kK1 = 1ull << 20;   // 1.0
constexpr unsigned long long kK2 = 1482910ull;   // round(sqrt(2) * 2^20)

struct GridPathArgs {
    const uint32_t* occ_bits;
    int words_per_map;
    const int32_t* task_to_map;
    const int32_t* starts;  // (y, x) as get_start_finish_coordinates returns them (generate_paths.py:81-87)
    const int32_t* goals;
    int B, H, W;
    unsigned long long* work;  // [gridDim.x][H * W]
    int32_t* queue;            // zeroed task counter
    int32_t* n_straight;       // [B], -1 = no path
    int32_t* n_diag;
    double* length;            // [B], NaN = no path
};

__device__ __forceinline__ unsigned long long make_key(unsigned a, unsigned b) {
    return ((a * kK1 + b * kK2) << 32) | ((unsigned long long)a << 16) | (unsigned long long)b;
}

__global__ void __launch_bounds__(kNT, 1) grid_path_kernel(const GridPathArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* grid = reinterpret_cast<uint32_t*>(smem_raw);                 // [words_per_map] (padded to 4)
    uint32_t* q = grid + ((a.words_per_map + 3) & ~3);                      // [3][kQCap]
    __shared__ int cnt[3], ovf[3], s_task, s_state;
    const int tid = threadIdx.x;
    const int cells = a.H * a.W;
    unsigned long long* dist = a.work + (size_t)blockIdx.x * cells;
    auto is_free = [&](int y, int x) -> bool {
        const int i = y * a.W + x;
        return (grid[i >> 5] >> (i & 31)) & 1u;
    };
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
        const int sy = a.starts[task * 2], sx = a.starts[task * 2 + 1], gy = a.goals[task * 2], gx = a.goals[task * 2 + 1];
        const int s = sy * a.W + sx, g = gy * a.W + gx;
        if (tid < 3) { cnt[tid] = 0; ovf[tid] = 0; }
        __syncthreads();
        if (tid == 0) { dist[s] = make_key(0, 0); q[0] = (uint32_t)s; cnt[0] = 1; s_state = 0; }
        __syncthreads();
        for (int k = 0; k < 4095; ++k) {
            const int cur = k % 3;
            const bool scan = ovf[cur] != 0;
            const int n = scan ? cells : min(cnt[cur], kQCap);
            for (int i = tid; i < n; i += kNT) {
                const int u = scan ? i : (int)q[cur * kQCap + i];
                const unsigned long long key = __ldcg(dist + u);  // labels are updated by L2 atomics: never read them through L1
                if (key == kInf || (int)(key >> 52) != k) continue;  // stale entry: the node moved to an earlier bucket
                const unsigned na0 = (unsigned)(key >> 16) & 0xffffu, nb0 = (unsigned)key & 0xffffu;
                const int y = u / a.W, x = u - y * a.W;
#pragma unroll
                for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                    for (int dx = -1; dx <= 1; ++dx) {
                        if (dy == 0 && dx == 0) continue;
                        const int yy = y + dy, xx = x + dx;
                        if (yy < 0 || yy >= a.H || xx < 0 || xx >= a.W || !is_free(yy, xx)) continue;
                        const bool diag = dy != 0 && dx != 0;
                        if (diag && (!is_free(yy, x) || !is_free(y, xx))) continue;  // no corner cutting
                        const unsigned long long nk = make_key(na0 + (diag ? 0u : 1u), nb0 + (diag ? 1u : 0u));
                        const int v = yy * a.W + xx;
                        if (nk < __ldcg(dist + v)) {  // cheap pre-test: most neighbours are already labelled better
                            const unsigned long long old = atomicMin(dist + v, nk);
                            if (nk < old) {
                                const int bkt = (int)(nk >> 52) % 3;
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
                const int n1 = (k + 1) % 3, n2 = (k + 2) % 3;
                cnt[cur] = 0; ovf[cur] = 0;
                if (gk != kInf && (int)(gk >> 52) <= k) s_state = 1;                                 // goal label final
                else if (cnt[n1] == 0 && cnt[n2] == 0 && !ovf[n1] && !ovf[n2]) s_state = 2;          // nothing left: no path
            }
            __syncthreads();
            if (s_state) break;
        }
        if (tid == 0) {
            const unsigned long long gk = __ldcg(dist + g);
            if (s_state == 1 && gk != kInf) {
                const int na = (int)((gk >> 16) & 0xffffu), nb = (int)(gk & 0xffffu);
                a.n_straight[task] = na; a.n_diag[task] = nb;
                a.length[task] = (double)na + (double)nb * 1.4142135623730951;
            } else {
                a.n_straight[task] = -1; a.n_diag[task] = -1;
                a.length[task] = nan("");
            }
        }
    }
}

}  // namespace

extern "C" int nrrt_grid_paths_2d(const uint32_t* occ_bits, int words_per_map, const int32_t* task_to_map, const int32_t* starts,
                                  const int32_t* goals, int B, int H, int W, void* work, int n_slots, int32_t* queue,
                                  int32_t* n_straight, int32_t* n_diag, double* length, void* stream) {
    NRRT_REQUIRE(occ_bits && task_to_map && starts && goals && work && queue && n_straight && n_diag && length, "null pointer");
    NRRT_REQUIRE(B >= 0 && H >= 1 && W >= 1 && H <= 2048 && W <= 2048 && words_per_map * 32 >= H * W, "bad grid shape");
    NRRT_REQUIRE(n_slots >= 1, "need at least one work slot");
    if (B == 0) return NRRT_OK;
    const size_t smem = (size_t)((words_per_map + 3) & ~3) * 4 + (size_t)3 * kQCap * 4;
    NRRT_REQUIRE(smem <= 227 * 1024, "grid too large for shared memory (%zu bytes)", smem);
    NRRT_CUDA_TRY(cudaFuncSetAttribute(grid_path_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GridPathArgs a;
    a.occ_bits = occ_bits; a.words_per_map = words_per_map; a.task_to_map = task_to_map; a.starts = starts; a.goals = goals;
    a.B = B; a.H = H; a.W = W; a.work = (unsigned long long*)work; a.queue = queue;
    a.n_straight = n_straight; a.n_diag = n_diag; a.length = length;
    const int grid = n_slots < B ? n_slots : B;
    grid_path_kernel<<<grid, kNT, smem, reinterpret_cast<cudaStream_t>(stream)>>>(a);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}
