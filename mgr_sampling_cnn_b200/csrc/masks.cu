// masks.cu — the reference's training-mask makers, batched on the device (HBM bound: one pass over the task).
//
// 2D  generate_paths.visualize_path (generate_paths.py:118-169), the lines that make the saved mask:
//       disks  = union of cv2.circle(path point, radius 10, filled)                   :129-131
//       blur   = cv2.GaussianBlur(disks, (33, 33), sigmaX = 3)  [8-bit fixed point]    :134
//       blur[path point] = 255                                                         :136-137
//       mask   = blur where occ_map > 240, else 0                                      :148,155
//     One CTA per 64-pixel column of tiles of a task (path bounding box taken once; unreachable tiles are zero-filled); per
//     64 x 64 tile the disks are rasterised straight into shared memory over the tile + blur halo
//     (they never exist in HBM), out-of-image halo cells are filled by BORDER_REFLECT_101, then the separable fixed-point
//     blur runs from shared memory (16-bit horizontal sums, 32-bit vertical sums, (v + 2^15) >> 16: OpenCV's
//     ufixedpoint16 / ufixedpoint32 arithmetic, integers only, so the result is bit-identical).  Algorithmic HBM bytes per
//     task: h*w (occupancy) + h*w (mask) + 8 * path points.
// 3D  ThreeD_enlarge_path.enlarge_the_path (ThreeD_enlarge_path.py:25-65): every free voxel within Chebyshev distance
//     d <= LAYERS-1 of a path voxel gets max(old, value[d]).  The Chebyshev distance transform is separable
//     (min over p of max(|dx|,|dy|,|dz|) = nested 1D min-max passes), so a CTA loads a 16^3 tile + halo of the path volume
//     as {0, INF} bytes and runs three 1D passes in shared memory.  Algorithmic HBM bytes per task: 3 * cells.
#include "nrrt_common.cuh"

#include <math.h>

namespace {

constexpr int kTile = 64;
constexpr int kMaxTaps = 65;      // ksize <= 65
constexpr int kMaxRadius = 31;

struct Mask2dArgs {
    const uint8_t* occ;
    const int32_t* task_to_map;
    const int32_t* paths;
    const int32_t* path_n;
    uint8_t* masks;
    int path_cap, B, h, w, radius, r_eff, free_above;
    uint16_t taps[kMaxTaps];      // taps[r_eff + i], i = -r_eff..r_eff
    int8_t half_width[kMaxRadius + 1];
};

__device__ __forceinline__ int reflect101(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

__global__ void __launch_bounds__(256) mask2d_kernel(const __grid_constant__ Mask2dArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int R = a.r_eff, RW = kTile + 2 * R, RS = (RW + 3) & ~3;       // region width, row stride of the disk tile
    uint8_t* disk = smem;                                                // [RW][RS]
    uint16_t* hsum = reinterpret_cast<uint16_t*>(smem + ((RW * RS + 15) & ~15));   // [RW][kTile]
    // A CTA owns one 64-pixel COLUMN of tiles of one task: it takes the bounding box of the path once and walks down the
    // column; tiles the box (grown by the circle radius and the blur reach) does not touch are zero-filled without looking at
    // the path again.  (One CTA per tile spent most of the kernel launching 262144 blocks that each rescanned the path.)
    const int task = blockIdx.y, tx0 = blockIdx.x * kTile;
    const int rx0 = tx0 - R;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_pts = min(a.path_n[task], a.path_cap);
    const int32_t* path = a.paths + (size_t)task * a.path_cap * 2;
    const uint8_t* occ = a.occ + (size_t)a.task_to_map[task] * a.h * a.w;
    uint8_t* mask = a.masks + (size_t)task * a.h * a.w;
    __shared__ int s_bb[4];
    if (tid == 0) { s_bb[0] = s_bb[2] = 0x7fffffff; s_bb[1] = s_bb[3] = -0x7fffffff; }
    __syncthreads();
    for (int p = tid; p < n_pts; p += 256) {
        atomicMin(&s_bb[0], path[2 * p]); atomicMax(&s_bb[1], path[2 * p]);
        atomicMin(&s_bb[2], path[2 * p + 1]); atomicMax(&s_bb[3], path[2 * p + 1]);
    }
    __syncthreads();
    const int reach = a.radius + R;
    const bool col_hit = n_pts > 0 && s_bb[2] - reach < tx0 + kTile && s_bb[3] + reach >= tx0;
    const int by0 = s_bb[0] - reach, by1 = s_bb[1] + reach;
    for (int ty0 = 0; ty0 < a.h; ty0 += kTile) {
        const int ry0 = ty0 - R;
        int touched = 0;
        __syncthreads();                                   // the previous tile is done with the shared-memory buffers
        if (col_hit && by0 < ty0 + kTile && by1 >= ty0) {   // else: the path cannot reach this tile (touched stays 0)
            for (int i = tid; i < RW * RS / 4; i += 256) reinterpret_cast<uint32_t*>(disk)[i] = 0u;
            __syncthreads();
            // 1. filled circles (cv::circle thickness < 0: row cy + dy spans cx +- half_width[|dy|]), clipped to the image
            for (int p = warp; p < n_pts; p += 8) {
                const int py = path[2 * p], px = path[2 * p + 1];
                if (py + a.radius < ry0 || py - a.radius >= ry0 + RW || px + a.radius < rx0 || px - a.radius >= rx0 + RW) continue;
                touched = 1;
                for (int dy = -a.radius + lane; dy <= a.radius; dy += 32) {
                    const int y = py + dy;
                    if (y < 0 || y >= a.h || y < ry0 || y >= ry0 + RW) continue;
                    const int hw = a.half_width[dy < 0 ? -dy : dy];
                    const int xa = max(max(px - hw, 0), rx0), xb = min(min(px + hw, a.w - 1), rx0 + RW - 1);
                    uint8_t* row = disk + (y - ry0) * RS - rx0;
                    for (int x = xa; x <= xb; ++x) row[x] = 255;
                }
            }
        }
        if (!__syncthreads_or(touched)) {   // no disk reaches this tile or its blur halo (most tiles of a task): the mask is 0 here
            for (int i = tid; i < kTile * kTile / 4; i += 256) {
                const int r = i / (kTile / 4), c = (i - r * (kTile / 4)) * 4, y = ty0 + r, x = tx0 + c;
                if (y >= a.h) continue;
                if (x + 3 < a.w && (a.w & 3) == 0) *reinterpret_cast<uint32_t*>(mask + (size_t)y * a.w + x) = 0u;
                else for (int j = 0; j < 4 && x + j < a.w; ++j) mask[(size_t)y * a.w + x + j] = 0;
            }
            continue;
        }
        // 2. BORDER_REFLECT_101 for halo cells outside the image
        for (int i = tid; i < RW * RW; i += 256) {
            const int r = i / RW, c = i - r * RW, y = ry0 + r, x = rx0 + c;
            if (y >= 0 && y < a.h && x >= 0 && x < a.w) continue;
            const int sy = reflect101(y, a.h) - ry0, sx = reflect101(x, a.w) - rx0;
            disk[r * RS + c] = (sy >= 0 && sy < RW && sx >= 0 && sx < RW) ? disk[sy * RS + sx] : 0;
        }
        __syncthreads();
        // 3. horizontal pass (8-bit x 8.8 fixed point -> 16 bit; the taps sum to 256 so 255 * 256 is the maximum)
        for (int i = tid; i < RW * kTile; i += 256) {
            const int r = i / kTile, c = i - r * kTile;
            const uint8_t* src = disk + r * RS + c;
            uint32_t s = 0;
            for (int k = 0; k <= 2 * R; ++k) s += (uint32_t)a.taps[k] * src[k];
            hsum[i] = (uint16_t)s;
        }
        __syncthreads();
        // 4. vertical pass, rounding, free-space mask; 4 pixels per thread
        for (int i = tid; i < kTile * kTile / 4; i += 256) {
            const int r = i / (kTile / 4), c = (i - r * (kTile / 4)) * 4, y = ty0 + r;
            if (y >= a.h) continue;
            uint32_t s[4] = {0, 0, 0, 0};
            for (int k = 0; k <= 2 * R; ++k) {
                const uint2 v = *reinterpret_cast<const uint2*>(hsum + (r + k) * kTile + c);
                const uint32_t t = a.taps[k];
                s[0] += t * (v.x & 0xffffu); s[1] += t * (v.x >> 16); s[2] += t * (v.y & 0xffffu); s[3] += t * (v.y >> 16);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int x = tx0 + c + j;
                if (x >= a.w) break;
                const uint32_t v = min((s[j] + 32768u) >> 16, 255u);
                mask[(size_t)y * a.w + x] = occ[(size_t)y * a.w + x] > a.free_above ? (uint8_t)v : (uint8_t)0;
            }
        }
        __syncthreads();
        // 5. the path pixels themselves are set to 255 after the blur (still under the free-space mask)
        for (int p = tid; p < n_pts; p += 256) {
            const int py = path[2 * p], px = path[2 * p + 1];
            if (py >= ty0 && py < ty0 + kTile && px >= tx0 && px < tx0 + kTile && py < a.h && px < a.w && py >= 0 && px >= 0)
                mask[(size_t)py * a.w + px] = occ[(size_t)py * a.w + px] > a.free_above ? 255 : 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
constexpr int kT3 = 16;

struct Enlarge3dArgs {
    const uint8_t* occ;
    const int32_t* task_to_map;
    const uint8_t* path_vols;
    uint8_t* out;
    int B, s0, s1, s2, layers;
    uint8_t values[8];
};

__global__ void __launch_bounds__(256) enlarge3d_kernel(const __grid_constant__ Enlarge3dArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int R = a.layers - 1, E = kT3 + 2 * R;
    uint8_t* d0 = smem;                       // [E][E][E]   (x, y, z) = array axes 0, 1, 2
    uint8_t* d1 = smem + E * E * E;           // [E][E][kT3] after the z pass, then reused
    const int tz = (a.s2 + kT3 - 1) / kT3, ty = (a.s1 + kT3 - 1) / kT3;
    const int task = blockIdx.y;
    int t = blockIdx.x;
    const int z0 = (t % tz) * kT3; t /= tz;
    const int y0 = (t % ty) * kT3; t /= ty;
    const int x0 = t * kT3;
    const size_t cells = (size_t)a.s0 * a.s1 * a.s2;
    const uint8_t* pv = a.path_vols + (size_t)task * cells;
    const uint8_t* occ = a.occ + (size_t)a.task_to_map[task] * cells;
    uint8_t* out = a.out + (size_t)task * cells;
    const int tid = threadIdx.x;
    // distance 0 on the path, INF elsewhere and outside the volume
    int touched = 0;
    // z rows of whole, aligned words (80^3 maps with LAYERS = 5): the tile + halo is read and written 4 / 16 voxels at a time
    const bool fast = (a.s2 & 15) == 0 && (R & 3) == 0 && (((uintptr_t)pv | (uintptr_t)occ | (uintptr_t)out) & 15) == 0;
    if (fast) {
        const int EW = E >> 2;
        uint32_t* d0w = reinterpret_cast<uint32_t*>(d0);
        for (int i = tid; i < E * E * EW; i += 256) {
            const int wi = i % EW, y = (i / EW) % E, x = i / (EW * E);
            const int gx = x0 - R + x, gy = y0 - R + y, gz = z0 - R + 4 * wi;
            uint32_t v = 0xffffffffu;
            if (gx >= 0 && gx < a.s0 && gy >= 0 && gy < a.s1 && gz >= 0 && gz < a.s2)
                v = ~__vcmpeq4(*reinterpret_cast<const uint32_t*>(pv + ((size_t)gx * a.s1 + gy) * a.s2 + gz), 0xffffffffu);
            d0w[i] = v;
            touched |= v != 0xffffffffu;
        }
    } else {
        for (int i = tid; i < E * E * E; i += 256) {
            const int z = i % E, y = (i / E) % E, x = i / (E * E);
            const int gx = x0 - R + x, gy = y0 - R + y, gz = z0 - R + z;
            uint8_t v = 255;
            if (gx >= 0 && gx < a.s0 && gy >= 0 && gy < a.s1 && gz >= 0 && gz < a.s2)
                v = pv[((size_t)gx * a.s1 + gy) * a.s2 + gz] == 255 ? 0 : 255;
            d0[i] = v;
            touched |= v == 0;
        }
    }
    if (!__syncthreads_or(touched)) {   // no path voxel within reach of this tile: the volume is copied through
        if (fast) {
            for (int i = tid; i < kT3 * kT3; i += 256) {         // one 16-voxel z row per thread
                const int gy = y0 + i % kT3, gx = x0 + i / kT3;
                if (gx >= a.s0 || gy >= a.s1) continue;
                const size_t g = ((size_t)gx * a.s1 + gy) * a.s2 + z0;
                *reinterpret_cast<uint4*>(out + g) = *reinterpret_cast<const uint4*>(pv + g);
            }
            return;
        }
        for (int i = tid; i < kT3 * kT3 * kT3; i += 256) {
            const int gz = z0 + i % kT3, gy = y0 + (i / kT3) % kT3, gx = x0 + i / (kT3 * kT3);
            if (gx >= a.s0 || gy >= a.s1 || gz >= a.s2) continue;
            const size_t g = ((size_t)gx * a.s1 + gy) * a.s2 + gz;
            out[g] = pv[g];
        }
        return;
    }
    if ((E & 3) == 0) {
        // Even reach (the reference's LAYERS = 5): four voxels of a z row per 32-bit word, byte-SIMD min / max.  A shifted window
        // of a row is a funnel shift of two neighbouring words; the y and x passes shift by whole rows, i.e. whole words.
        const int EW = E >> 2;                                   // words per z row of d0
        const uint32_t* d0w = reinterpret_cast<const uint32_t*>(d0);
        uint32_t* d1w = reinterpret_cast<uint32_t*>(d1);
        for (int i = tid; i < E * E * 4; i += 256) {             // z pass: d1[x][y][16 z]
            const int j = i & 3, row = i >> 2;
            const uint32_t* rw = d0w + row * EW;
            uint32_t m = 0xffffffffu;
            for (int dz = -R; dz <= R; ++dz) {
                const int o = 4 * j + R + dz, wi = o >> 2;
                const uint32_t lo = rw[wi], hi = wi + 1 < EW ? rw[wi + 1] : 0u;
                const uint32_t wv = __funnelshift_r(lo, hi, (o & 3) * 8);
                m = __vminu4(m, __vmaxu4(wv, (uint32_t)(dz < 0 ? -dz : dz) * 0x01010101u));
            }
            d1w[i] = m;
        }
        __syncthreads();
        uint32_t* d0o = reinterpret_cast<uint32_t*>(d0);
        for (int i = tid; i < E * kT3 * 4; i += 256) {           // y pass into d0 as [E][16][16]
            const int j = i & 3, y = (i >> 2) % kT3, x = i / (4 * kT3);
            const uint32_t* src = d1w + (x * E + y + R) * 4 + j;
            uint32_t m = 0xffffffffu;
            for (int dy = -R; dy <= R; ++dy) m = __vminu4(m, __vmaxu4(src[dy * 4], (uint32_t)(dy < 0 ? -dy : dy) * 0x01010101u));
            d0o[i] = m;
        }
        __syncthreads();
        for (int i = tid; i < kT3 * kT3 * 4; i += 256) {         // x pass + paint
            const int j = i & 3, y = (i >> 2) % kT3, x = i / (4 * kT3);
            const int gx = x0 + x, gy = y0 + y;
            if (gx >= a.s0 || gy >= a.s1) continue;
            const uint32_t* src = d0o + ((x + R) * kT3 + y) * 4 + j;
            uint32_t m = 0xffffffffu;
            for (int dx = -R; dx <= R; ++dx) m = __vminu4(m, __vmaxu4(src[dx * kT3 * 4], (uint32_t)(dx < 0 ? -dx : dx) * 0x01010101u));
            if (fast) {
                const size_t g = ((size_t)gx * a.s1 + gy) * a.s2 + z0 + 4 * j;
                const uint32_t pw = *reinterpret_cast<const uint32_t*>(pv + g), ow = *reinterpret_cast<const uint32_t*>(occ + g);
                uint32_t res = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int mk = (m >> (8 * k)) & 0xff;
                    uint32_t v = (pw >> (8 * k)) & 0xff;
                    if (mk < a.layers && ((ow >> (8 * k)) & 0xff) != 255 && a.values[mk] > v) v = a.values[mk];
                    res |= v << (8 * k);
                }
                *reinterpret_cast<uint32_t*>(out + g) = res;
                continue;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int gz = z0 + 4 * j + k;
                if (gz >= a.s2) break;
                const int mk = (m >> (8 * k)) & 0xff;
                const size_t g = ((size_t)gx * a.s1 + gy) * a.s2 + gz;
                uint8_t v = pv[g];
                if (mk < a.layers && occ[g] != 255 && a.values[mk] > v) v = a.values[mk];
                out[g] = v;
            }
        }
        return;
    }
    // z pass: d1[x][y][z] = min over dz of max(d0[x][y][z + dz], |dz|), z in the tile
    for (int i = tid; i < E * E * kT3; i += 256) {
        const int z = i % kT3, xy = i / kT3;
        const uint8_t* src = d0 + xy * E + z + R;
        int m = 255;
        for (int dz = -R; dz <= R; ++dz) m = min(m, max((int)src[dz], dz < 0 ? -dz : dz));
        d1[i] = (uint8_t)m;
    }
    __syncthreads();
    // y pass into d0 (as [E][kT3][kT3])
    for (int i = tid; i < E * kT3 * kT3; i += 256) {
        const int z = i % kT3, y = (i / kT3) % kT3, x = i / (kT3 * kT3);
        const uint8_t* src = d1 + (x * E + y + R) * kT3 + z;
        int m = 255;
        for (int dy = -R; dy <= R; ++dy) m = min(m, max((int)src[dy * kT3], dy < 0 ? -dy : dy));
        d0[i] = (uint8_t)m;
    }
    __syncthreads();
    // x pass + paint
    for (int i = tid; i < kT3 * kT3 * kT3; i += 256) {
        const int z = i % kT3, y = (i / kT3) % kT3, x = i / (kT3 * kT3);
        const int gx = x0 + x, gy = y0 + y, gz = z0 + z;
        if (gx >= a.s0 || gy >= a.s1 || gz >= a.s2) continue;
        const uint8_t* src = d0 + ((x + R) * kT3 + y) * kT3 + z;
        int m = 255;
        for (int dx = -R; dx <= R; ++dx) m = min(m, max((int)src[dx * kT3 * kT3], dx < 0 ? -dx : dx));
        const size_t g = ((size_t)gx * a.s1 + gy) * a.s2 + gz;
        uint8_t v = pv[g];
        if (m < a.layers && occ[g] != 255 && a.values[m] > v) v = a.values[m];
        out[g] = v;
    }
}

__global__ void paths_to_volume_kernel(const int32_t* paths, const int32_t* path_n, int path_cap, int B, int s0, int s1, int s2,
                                       uint8_t* vols) {
    const int task = blockIdx.x;
    const int n = min(path_n[task], path_cap);
    const size_t cells = (size_t)s0 * s1 * s2;
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
        const int32_t* q = paths + ((size_t)task * path_cap + p) * 3;
        const int x = q[0], y = q[1], z = q[2];
        if (x >= 0 && x < s0 && y >= 0 && y < s1 && z >= 0 && z < s2) vols[task * cells + ((size_t)x * s1 + y) * s2 + z] = 255;
    }
}

// cv::circle(thickness < 0, LINE_8, shift 0) -> Circle() of imgproc/drawing.cpp: midpoint walk
void disk_half_widths(int radius, int8_t* hw) {
    for (int i = 0; i <= radius; ++i) hw[i] = -1;
    int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
    while (dx >= dy) {
        if (dx > hw[dy]) hw[dy] = (int8_t)dx;
        if (dy > hw[dx]) hw[dx] = (int8_t)dy;
        ++dy;
        err += plus;
        plus += 2;
        const int mask = (err <= 0) - 1;
        err -= minus & mask;
        dx += mask;
        minus -= mask & 2;
    }
}

// getGaussianKernelFixedPoint_ED of imgproc/smooth.dispatch.cpp: 8.8 fixed-point taps, rounding error diffused towards the centre
void gaussian_taps_fixed(int n, double sigma, int64_t* res) {
    double k[kMaxTaps], sum = 0;
    const double scale = -0.5 / (sigma * sigma);
    for (int i = 0; i < n; ++i) {
        const double x = i - (n - 1) / 2;
        k[i] = exp(scale * x * x);
        sum += k[i];
    }
    double err = 0;
    int64_t s = 0;
    for (int i = 0; i < n / 2; ++i) {
        const double adj = k[i] / sum * 256.0 + err;
        const int64_t v0 = (int64_t)nearbyint(adj);
        err = adj - (double)v0;
        res[i] = res[n - 1 - i] = v0;
        s += v0;
    }
    res[n / 2] = 256 - 2 * s;
}

}  // namespace

extern "C" int nrrt_path_masks_2d(const uint8_t* occ_maps, const int32_t* task_to_map, const int32_t* paths,
                                  const int32_t* path_n, int path_cap, int B, int h, int w, int radius, int ksize, double sigma,
                                  int free_above, uint8_t* masks, void* stream) {
    NRRT_REQUIRE(B >= 0 && path_cap >= 1, "bad batch sizes");
    NRRT_REQUIRE(radius >= 0 && radius <= kMaxRadius, "circle radius must be in [0, %d]", kMaxRadius);
    NRRT_REQUIRE(ksize >= 1 && (ksize & 1) && ksize <= kMaxTaps && sigma > 0, "ksize must be odd and <= %d, sigma > 0", kMaxTaps);
    NRRT_REQUIRE(h > ksize / 2 && w > ksize / 2 && h <= 32768 && w <= 32768, "map sides must exceed ksize / 2");
    if (B == 0) return NRRT_OK;
    NRRT_REQUIRE(occ_maps && task_to_map && paths && path_n && masks, "null pointer");
    Mask2dArgs a;
    a.occ = occ_maps; a.task_to_map = task_to_map; a.paths = paths; a.path_n = path_n; a.masks = masks;
    a.path_cap = path_cap; a.B = B; a.h = h; a.w = w; a.radius = radius; a.free_above = free_above;
    disk_half_widths(radius, a.half_width);
    int64_t taps[kMaxTaps];
    gaussian_taps_fixed(ksize, sigma, taps);
    int r_eff = 0;
    for (int i = 0; i < ksize; ++i) {
        NRRT_REQUIRE(taps[i] >= 0 && taps[i] <= 256, "fixed-point Gaussian tap out of range");
        if (taps[i] != 0) r_eff = r_eff > abs(i - ksize / 2) ? r_eff : abs(i - ksize / 2);
    }
    a.r_eff = r_eff;
    for (int i = -r_eff; i <= r_eff; ++i) a.taps[r_eff + i] = (uint16_t)taps[ksize / 2 + i];
    const int RW = kTile + 2 * r_eff, RS = (RW + 3) & ~3;
    const size_t smem = ((RW * RS + 15) & ~15) + (size_t)RW * kTile * 2;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (smem > 48 * 1024) NRRT_CUDA_TRY(cudaFuncSetAttribute(mask2d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NRRT_REQUIRE(B <= 65535, "at most 65535 tasks per call");
    const dim3 grid((w + kTile - 1) / kTile, B);
    mask2d_kernel<<<grid, 256, smem, st>>>(a);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_mask_tables_host(int radius, int ksize, double sigma, int32_t* half_width, int32_t* taps) {
    NRRT_REQUIRE(radius >= 0 && radius <= kMaxRadius && ksize >= 1 && (ksize & 1) && ksize <= kMaxTaps && sigma > 0, "bad arguments");
    NRRT_REQUIRE(half_width && taps, "null pointer");
    int8_t hw[kMaxRadius + 1];
    int64_t t[kMaxTaps];
    disk_half_widths(radius, hw);
    gaussian_taps_fixed(ksize, sigma, t);
    for (int i = 0; i <= radius; ++i) half_width[i] = hw[i];
    for (int i = 0; i < ksize; ++i) taps[i] = (int32_t)t[i];
    return NRRT_OK;
}

extern "C" int nrrt_paths_to_volume(const int32_t* paths, const int32_t* path_n, int path_cap, int B, int s0, int s1, int s2,
                                    uint8_t* vols, void* stream) {
    NRRT_REQUIRE(B >= 0 && path_cap >= 1 && s0 >= 1 && s1 >= 1 && s2 >= 1, "bad sizes");
    if (B == 0) return NRRT_OK;
    NRRT_REQUIRE(paths && path_n && vols, "null pointer");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    NRRT_CUDA_TRY(cudaMemsetAsync(vols, 0, (size_t)B * s0 * s1 * s2, st));
    paths_to_volume_kernel<<<B, 128, 0, st>>>(paths, path_n, path_cap, B, s0, s1, s2, vols);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}

extern "C" int nrrt_enlarge_paths_3d(const uint8_t* occ_maps, const int32_t* task_to_map, const uint8_t* path_vols, int B, int s0,
                                     int s1, int s2, int layers, const uint8_t* values_host, uint8_t* out, void* stream) {
    NRRT_REQUIRE(B >= 0 && s0 >= 1 && s1 >= 1 && s2 >= 1, "bad sizes");
    NRRT_REQUIRE(layers >= 1 && layers <= 6, "layers must be in [1, 6]");
    if (B == 0) return NRRT_OK;
    NRRT_REQUIRE(occ_maps && task_to_map && path_vols && values_host && out, "null pointer");
    NRRT_REQUIRE(out != path_vols, "out must not alias path_vols (tiles read their neighbours' halo)");
    NRRT_REQUIRE(B <= 65535, "at most 65535 tasks per call");
    Enlarge3dArgs a;
    a.occ = occ_maps; a.task_to_map = task_to_map; a.path_vols = path_vols; a.out = out;
    a.B = B; a.s0 = s0; a.s1 = s1; a.s2 = s2; a.layers = layers;
    for (int i = 0; i < 8; ++i) a.values[i] = i < layers ? values_host[i] : 0;
    const int E = kT3 + 2 * (layers - 1);
    const size_t smem = (size_t)E * E * E + (size_t)E * E * kT3;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (smem > 48 * 1024) NRRT_CUDA_TRY(cudaFuncSetAttribute(enlarge3d_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = ((s0 + kT3 - 1) / kT3) * ((s1 + kT3 - 1) / kT3) * ((s2 + kT3 - 1) / kT3);
    enlarge3d_kernel<<<dim3(tiles, B), 256, smem, st>>>(a);
    NRRT_CUDA_TRY(cudaGetLastError());
    return NRRT_OK;
}
