// generate.cu — on-device workload generation: the reference's map and task generators with CPython's
// `random` reproduced bit for bit, so that `random.seed(s)` on the host and seed s here give the same maps and tasks.
//
//   generate_maps.create_map / add_obstacle            generate_maps.py:7-44            (2D: 0 = obstacle, 255 = free)
//   ThreeD_generate_maps.create_map / add_obstacle     ThreeD_generate_maps.py:5-41     (3D: 255 = obstacle, 0 = free)
//   generate_tasks.draw_start_and_finish               generate_tasks.py:27-46          (free start, free goal >= 100 px away)
//   ThreeD_generate_tasks.draw_start_and_finish        ThreeD_generate_tasks.py:25-48   (>= 20 voxels away)
//
// CPython's generator: MT19937 seeded by random.seed(int) = init_by_array(32-bit words of |seed|); randint(a, b) =
// a + _randbelow(b - a + 1); _randbelow(n) = getrandbits(n.bit_length()) redrawn while >= n; getrandbits(k <= 32) = the
// top k bits of one 32-bit output (Lib/random.py, Modules/_randommodule.c).  The draw ORDER is the reference's: every
// call site below is annotated with the line it mirrors.
//
// One CTA per map.  The generator is a serial recurrence, so one thread runs it (state in shared memory) and emits the
// map's boxes / the task candidates; all threads of the CTA then rasterise the boxes (coalesced 16-byte stores).
// Task candidates: the reference pipeline deletes tasks its A* cannot solve (generate_paths.py:111-113), which does not
// touch the random stream -- so a map's candidates are drawn back to back (the stream does not depend on which are kept),
// reachability is decided for all of them at once by nrrt_grid_paths, and nrrt_select_tasks keeps the first `per_map`
// reachable ones per map, exactly the tasks the sequential host loop (mgr_sampling_cnn_b200/workload.py) keeps.
#include "nrrt_common.cuh"

namespace {

struct Mt {  // MT19937, state in shared memory, driven by one thread
    uint32_t* mt;
    int idx;

    __device__ void seed(unsigned long long s) {  // random.seed(int): init_by_array(key), key = 32-bit words of s
        const uint32_t key[2] = {(uint32_t)s, (uint32_t)(s >> 32)};
        const int klen = key[1] ? 2 : 1;
        mt[0] = 19650218u;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        int i = 1, j = 0;
        for (int k = 624; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
            ++i; ++j;
            if (i >= 624) { mt[0] = mt[623]; i = 1; }
            if (j >= klen) j = 0;
        }
        for (int k = 623; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
            ++i;
            if (i >= 624) { mt[0] = mt[623]; i = 1; }
        }
        mt[0] = 0x80000000u;
        idx = 624;
    }
    __device__ uint32_t next() {
        if (idx >= 624) {
            for (int k = 0; k < 624; ++k) {
                const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
                mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    __device__ int randint(int a, int b) {  // random.randint(a, b), b - a + 1 < 2^31
        const uint32_t n = (uint32_t)(b - a + 1);
        const int k = 32 - __clz(n);  // n.bit_length()
        uint32_t r = next() >> (32 - k);
        while (r >= n) r = next() >> (32 - k);
        return a + (int)r;
    }
};

constexpr int kMaxBoxes = 40;  // 2D: <= 7 + 7 + 3 + 15; 3D: <= 3 + 3 + 3 + 3 + 15

struct Box { int lo[3], hi[3]; };  // half-open [lo, hi) per axis, already clipped to the map

// ---- maps --------------------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(256) generate_maps_kernel(const long long* __restrict__ seeds, int n_maps, int s0, int s1, int s2,
                                                            uint8_t* __restrict__ occ) {
    __shared__ uint32_t state[624];
    __shared__ Box boxes[kMaxBoxes];
    __shared__ int n_boxes;
    const int m = blockIdx.x;
    if (m >= n_maps) return;
    if (threadIdx.x == 0) {
        Mt g{state, 624};
        g.seed((unsigned long long)seeds[m]);
        int nb = 0;
        if (DIM == 2) {
            // create_map(width, height): occ_map is (height, width) = (s0, s1); the placement is drawn from 0..512 whatever
            // the map size (generate_maps.py:35-36)
            const int spec[4][6] = {{2, 7, 75, 350, 25, 75}, {2, 7, 25, 75, 75, 350}, {1, 3, 75, 175, 75, 175}, {5, 15, 25, 125, 25, 125}};
            for (int t = 0; t < 4; ++t) {
                const int count = g.randint(spec[t][0], spec[t][1]);             // for i in range(0, random.randint(..)):
                for (int q = 0; q < count; ++q) {
                    const int xs = g.randint(0, 512);                            //   x_start = random.randint(0, 512)
                    const int ys = g.randint(0, 512);                            //   y_start = random.randint(0, 512)
                    const int xe = g.randint(xs + spec[t][2], xs + spec[t][3]);  //   x_end = random.randint(x_start + x_min, x_start + x_max)
                    const int ye = g.randint(ys + spec[t][4], ys + spec[t][5]);  //   y_end = ...
                    // cv2.rectangle(pt1=(xs, ys), pt2=(xe, ye), thickness=-1): the inclusive box, clipped to the image
                    Box b;
                    b.lo[0] = min(ys, s0); b.hi[0] = min(ye + 1, s0);
                    b.lo[1] = min(xs, s1); b.hi[1] = min(xe + 1, s1);
                    b.lo[2] = 0; b.hi[2] = 1;
                    if (nb < kMaxBoxes) boxes[nb++] = b;
                }
            }
        } else {
            // create_map(width, height, depth): occ_map = zeros((height, width, depth)) = (s0, s1, s2); axes x, y, z = 0, 1, 2
            const int spec[5][8] = {{1, 3, 20, 55, 10, 20, 10, 20}, {1, 3, 10, 20, 20, 55, 10, 20}, {1, 3, 10, 20, 10, 20, 20, 55},
                                    {1, 3, 20, 30, 20, 30, 20, 30}, {5, 15, 10, 20, 10, 20, 10, 20}};
            for (int t = 0; t < 5; ++t) {
                const int count = g.randint(spec[t][0], spec[t][1]);
                for (int q = 0; q < count; ++q) {
                    const int xs = g.randint(0, s0 - 1);                         //   x_start = random.randint(0, occ_map.shape[0] - 1)
                    const int ys = g.randint(0, s1 - 1);
                    const int zs = g.randint(0, s2 - 1);
                    const int xe = g.randint(xs + spec[t][2], xs + spec[t][3]);
                    const int ye = g.randint(ys + spec[t][4], ys + spec[t][5]);
                    const int ze = g.randint(zs + spec[t][6], zs + spec[t][7]);
                    Box b;  // occ_map[x_start:x_end, y_start:y_end, z_start:z_end] = 255 (exclusive ends, clipped by slicing)
                    b.lo[0] = xs; b.hi[0] = min(xe, s0);
                    b.lo[1] = ys; b.hi[1] = min(ye, s1);
                    b.lo[2] = zs; b.hi[2] = min(ze, s2);
                    if (nb < kMaxBoxes) boxes[nb++] = b;
                }
            }
        }
        n_boxes = nb;
    }
    __syncthreads();
    const int nb = n_boxes;
    const long long cells = (long long)s0 * s1 * s2;
    uint8_t* out = occ + (size_t)m * cells;
    const uint8_t inside = DIM == 2 ? 0 : 255, outside = DIM == 2 ? 255 : 0;
    // 16 cells of one innermost row per step (the innermost extent is a multiple of 16 for the supported sizes, else bytes)
    const int inner = DIM == 2 ? s1 : s2;
    const bool vec = inner % 16 == 0 && ((uintptr_t)out & 15) == 0;
    const long long steps = vec ? cells / 16 : cells;
    for (long long q = threadIdx.x; q < steps; q += blockDim.x) {
        const long long c0 = vec ? q * 16 : q;
        const int i2 = (int)(c0 % inner);
        long long r = c0 / inner;
        const int i1 = DIM == 2 ? 0 : (int)(r % s1);
        const int i0 = DIM == 2 ? (int)r : (int)(r / s1);
        uint32_t mask = 0;  // bit e: cell i2 + e is inside a box (or on the 2D frame)
        const int span = vec ? 16 : 1;
        if (DIM == 2) {
            if (i0 == 0 || i0 == s0 - 1) mask = 0xffffu;
            if (i2 == 0) mask |= 1u;
            if (i2 + span >= s1) mask |= 1u << (s1 - 1 - i2);
        }
        for (int b = 0; b < nb; ++b) {
            const Box& bx = boxes[b];
            const int a0 = DIM == 2 ? i0 : i0, a1 = i1;
            if (a0 < bx.lo[0] || a0 >= bx.hi[0]) continue;
            if (DIM == 3 && (a1 < bx.lo[1] || a1 >= bx.hi[1])) continue;
            const int lo = (DIM == 2 ? bx.lo[1] : bx.lo[2]) - i2, hi = (DIM == 2 ? bx.hi[1] : bx.hi[2]) - i2;
            const int l = max(lo, 0), h = min(hi, span);
            if (h > l) mask |= ((1u << (h - l)) - 1u) << l;
        }
        if (vec) {
            uint32_t wds[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t v = 0;
#pragma unroll
                for (int e = 0; e < 4; ++e) v |= (uint32_t)(((mask >> (4 * k + e)) & 1u) ? inside : outside) << (8 * e);
                wds[k] = v;
            }
            *reinterpret_cast<uint4*>(out + c0) = make_uint4(wds[0], wds[1], wds[2], wds[3]);
        } else {
            out[c0] = (mask & 1u) ? inside : outside;
        }
    }
}

// ---- task candidates ---------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(32) task_candidates_kernel(const long long* __restrict__ seeds, int n_maps, int s0, int s1, int s2,
                                                             const uint8_t* __restrict__ occ, int n_cand, int max_draws,
                                                             int32_t* __restrict__ starts, int32_t* __restrict__ goals,
                                                             int32_t* __restrict__ n_done) {
    __shared__ uint32_t state[624];
    const int m = blockIdx.x;
    if (m >= n_maps || threadIdx.x != 0) return;
    Mt g{state, 624};
    g.seed((unsigned long long)seeds[m]);
    const uint8_t* map = occ + (size_t)m * s0 * s1 * s2;
    int done = 0;
    long long budget = max_draws;  // a map without a free cell (or without two free cells far enough apart) would spin forever
    for (; done < n_cand && budget > 0; ++done) {
        int32_t* so = starts + ((size_t)m * n_cand + done) * DIM;
        int32_t* go = goals + ((size_t)m * n_cand + done) * DIM;
        if (DIM == 2) {
            // dimensions = occ_map.shape; x is drawn from dimensions[0] and y from dimensions[1] (generate_tasks.py:33-34)
            int xs = 0, ys = 0;
            while (map[(size_t)ys * s1 + xs] == 0 && budget > 0) {              // while occ_map[y_start, x_start] == 0:
                xs = g.randint(0, s0 - 1);
                ys = g.randint(0, s1 - 1);
                budget -= 2;
            }
            int xf = 0, yf = 0;
            double dst = 0.0;
            while ((map[(size_t)yf * s1 + xf] == 0 || dst < 100.0) && budget > 0) {  // while occ_map[yf, xf] == 0 or dst < 100.0:
                xf = g.randint(0, s0 - 1);
                yf = g.randint(0, s1 - 1);
                budget -= 2;
                const int dx = xs - xf, dy = ys - yf;
                dst = sqrt((double)(dx * dx + dy * dy));                         // distance.euclidean on ints: exact squares
            }
            so[0] = ys; so[1] = xs;   // planner order (y, x)
            go[0] = yf; go[1] = xf;
        } else {
            auto at = [&](int x, int y, int z) { return map[((size_t)x * s1 + y) * s2 + z]; };
            int xs = g.randint(0, s0 - 1), ys = g.randint(0, s1 - 1), zs = g.randint(0, s2 - 1);
            budget -= 3;
            while (at(xs, ys, zs) == 255 && budget > 0) {
                xs = g.randint(0, s0 - 1); ys = g.randint(0, s1 - 1); zs = g.randint(0, s2 - 1);
                budget -= 3;
            }
            int xf = g.randint(0, s0 - 1), yf = g.randint(0, s1 - 1), zf = g.randint(0, s2 - 1);
            budget -= 3;
            double dst = 0.0;
            while ((at(xf, yf, zf) == 255 || dst < 20.0) && budget > 0) {
                xf = g.randint(0, s0 - 1); yf = g.randint(0, s1 - 1); zf = g.randint(0, s2 - 1);
                budget -= 3;
                const int dx = xs - xf, dy = ys - yf, dz = zs - zf;
                dst = sqrt((double)(dx * dx + dy * dy + dz * dz));
            }
            so[0] = xs; so[1] = ys; so[2] = zs;
            go[0] = xf; go[1] = yf; go[2] = zf;
        }
    }
    n_done[m] = budget > 0 ? done : -1;
}

// ---- selection: the first `per_map` reachable candidates of every map, in candidate order ------------------------
template <int DIM>
__global__ void select_tasks_kernel(int n_maps, int n_cand, int per_map, int n_tasks, const int32_t* __restrict__ c_starts,
                                    const int32_t* __restrict__ c_goals, const int32_t* __restrict__ reach_steps,
                                    int32_t* __restrict__ task_to_map, int32_t* __restrict__ starts, int32_t* __restrict__ goals,
                                    float* __restrict__ coords, int32_t* __restrict__ redrawn, int32_t* __restrict__ short_maps) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_maps) return;
    const int want = min(per_map, n_tasks - m * per_map);  // the last map may be truncated
    int kept = 0, skipped = 0;
    for (int c = 0; c < n_cand && kept < want; ++c) {
        const size_t q = (size_t)m * n_cand + c;
        if (reach_steps && reach_steps[q * 3] < 0) { ++skipped; continue; }  // no path: the reference deletes the task
        const int t = m * per_map + kept;
        task_to_map[t] = m;
        for (int k = 0; k < DIM; ++k) { starts[t * DIM + k] = c_starts[q * DIM + k]; goals[t * DIM + k] = c_goals[q * DIM + k]; }
        // the CNN's coords input: [[sx, sy], [fx, fy]] in 2D (train.py:57-60), [[sx, sy, sz], [fx, fy, fz]] in 3D
        float* co = coords + (size_t)t * 2 * DIM;
        if (DIM == 2) {
            co[0] = (float)c_starts[q * 2 + 1]; co[1] = (float)c_starts[q * 2];
            co[2] = (float)c_goals[q * 2 + 1]; co[3] = (float)c_goals[q * 2];
        } else {
            for (int k = 0; k < 3; ++k) { co[k] = (float)c_starts[q * 3 + k]; co[3 + k] = (float)c_goals[q * 3 + k]; }
        }
        ++kept;
    }
    if (skipped) atomicAdd(redrawn, skipped);
    if (kept < want) atomicAdd(short_maps, 1);  // not enough reachable candidates: the caller retries with more
}

}  // namespace

extern "C" int nrrt_generate_maps(int dims, const int64_t* seeds, int n_maps, int s0, int s1, int s2, uint8_t* occ_maps,
                                  void* stream) {
    NRRT_REQUIRE(dims == 2 || dims == 3, "dims must be 2 or 3");
    if (dims == 2) s2 = 1;
    NRRT_REQUIRE(n_maps >= 0 && s0 >= 2 && s1 >= 2 && s2 >= 1 && s0 <= 32767 && s1 <= 32767 && s2 <= 32767, "bad map shape");
    if (n_maps == 0) return NRRT_OK;
    NRRT_REQUIRE(seeds && occ_maps, "null pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long* sd = reinterpret_cast<const long long*>(seeds);
    if (dims == 2) generate_maps_kernel<2><<<n_maps, 256, 0, st>>>(sd, n_maps, s0, s1, 1, occ_maps);
    else generate_maps_kernel<3><<<n_maps, 256, 0, st>>>(sd, n_maps, s0, s1, s2, occ_maps);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_generate_task_candidates(int dims, const int64_t* seeds, int n_maps, int s0, int s1, int s2,
                                             const uint8_t* occ_maps, int n_cand, int32_t* starts, int32_t* goals,
                                             int32_t* n_done, void* stream) {
    NRRT_REQUIRE(dims == 2 || dims == 3, "dims must be 2 or 3");
    if (dims == 2) s2 = 1;
    NRRT_REQUIRE(n_maps >= 0 && n_cand >= 1 && s0 >= 2 && s1 >= 2 && s2 >= 1, "bad sizes");
    NRRT_REQUIRE(dims == 3 || s0 == s1, "2D task drawing indexes occ_map[y, x] with x < shape[0], y < shape[1] (generate_tasks.py:33-35): square maps only");
    if (n_maps == 0) return NRRT_OK;
    NRRT_REQUIRE(seeds && occ_maps && starts && goals && n_done, "null pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const long long* sd = reinterpret_cast<const long long*>(seeds);
    const int max_draws = 4000000;
    if (dims == 2) task_candidates_kernel<2><<<n_maps, 32, 0, st>>>(sd, n_maps, s0, s1, 1, occ_maps, n_cand, max_draws, starts, goals, n_done);
    else task_candidates_kernel<3><<<n_maps, 32, 0, st>>>(sd, n_maps, s0, s1, s2, occ_maps, n_cand, max_draws, starts, goals, n_done);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_select_tasks(int dims, int n_maps, int n_cand, int per_map, int n_tasks, const int32_t* cand_starts,
                                 const int32_t* cand_goals, const int32_t* reach_steps, int32_t* task_to_map, int32_t* starts,
                                 int32_t* goals, float* coords, int32_t* redrawn, int32_t* short_maps, void* stream) {
    NRRT_REQUIRE(dims == 2 || dims == 3, "dims must be 2 or 3");
    NRRT_REQUIRE(n_maps >= 0 && n_cand >= 1 && per_map >= 1 && n_tasks >= 0 && n_tasks <= n_maps * per_map &&
                     n_tasks > (n_maps - 1) * per_map - (n_maps == 0), "bad sizes");
    if (n_maps == 0) return NRRT_OK;
    NRRT_REQUIRE(cand_starts && cand_goals && task_to_map && starts && goals && coords && redrawn && short_maps, "null pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    NRRT_CUDA_TRY(cudaMemsetAsync(redrawn, 0, sizeof(int32_t), st));
    NRRT_CUDA_TRY(cudaMemsetAsync(short_maps, 0, sizeof(int32_t), st));
    const int nt = 128, nb = (n_maps + nt - 1) / nt;
    if (dims == 2) select_tasks_kernel<2><<<nb, nt, 0, st>>>(n_maps, n_cand, per_map, n_tasks, cand_starts, cand_goals, reach_steps, task_to_map, starts, goals, coords, redrawn, short_maps);
    else select_tasks_kernel<3><<<nb, nt, 0, st>>>(n_maps, n_cand, per_map, n_tasks, cand_starts, cand_goals, reach_steps, task_to_map, starts, goals, coords, redrawn, short_maps);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}
