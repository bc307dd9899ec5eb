// dh_kernels.cu — hand-written sm_100a kernels for depthhead's Hough-forest prediction path.
//
// Pipeline per chunk of frames (reference: src/hough/prediction.rs:421-753):
//   K1  sat_rows / sat_cols      u16 depth -> u32 summed-area table (wrap-around exact, see below)
//   K2  traverse_kernel          TMA-staged SAT tile in shared memory; one thread per patch x tree
//                                walks root->leaf (houghforest.rs:185-193, types.rs:317-339)
//   K3a gate_kernel              ordered f64 prob sum, 0.7 gate, back-projection, hit list
//                                (prediction.rs:551-554,582-595)
//   K3b coarse_vote_kernel       coarse seed grids + arg-max seeds (prediction.rs:601-747)
//   K4a plan/clear/insert        per-frame open-addressing hash accumulators = SparseArray3D<u32>
//                                (meanshift.rs:14-68), restricted to the mean-shift reach of the seed
//   K4b meanshift_kernel         meanshift.rs:328-407 in reference accumulation order
//   K5  leaf_gate_kernel         estimate_mean_cov traces + valtoadd per leaf, once per model
//                                (meancov_estimation.rs:359-378, prediction.rs:594-600,643)
//
// Exactness rules: float->int is truncation (cvt.rzi, saturating, NaN->0 ==
// Rust `as`); no FMA contraction anywhere a float feeds a truncation or comparison (explicit
// __f*_rn / __d*_rn intrinsics, and the file is compiled with -fmad=false); vote sums are u32
// atomics (order independent, wrap like a release build).
//
// Summed-area table: S[y][x] = sum of pixels with x'<x, y'<y, in u32 with wrap-around.  Any
// rectangle whose true sum is < 2^32 (255*255*65535 < 2^32) is recovered exactly by the four-tap
// difference in modular arithmetic, and (double)sum/(double)count then equals the reference's
// naive loop bit for bit (types.rs:338).
#include "dh_kernels.cuh"

#include <cstdio>

namespace dh {

namespace dev {

constexpr unsigned long long kEmptyKey = ~0ull;
constexpr long long kKeyBias = 1ll << 20;

// ---------------------------------------------------------------- small device helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA: 3-D tiled bulk tensor load global -> shared, completion signalled on an mbarrier.
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return k;
}

// Key of an accumulator cell relative to the mean-shift seed; false if outside the stored reach.
__device__ __forceinline__ bool make_key(int x, int y, int z, const int32_t* seed, long long reach,
                                         unsigned long long* key) {
    const long long dx = (long long)x - seed[0], dy = (long long)y - seed[1], dz = (long long)z - seed[2];
    if (dx < -reach || dx > reach || dy < -reach || dy > reach || dz < -reach || dz > reach) return false;
    *key = (unsigned long long)(dx + kKeyBias) | ((unsigned long long)(dy + kKeyBias) << 21) |
           ((unsigned long long)(dz + kKeyBias) << 42);
    return true;
}

__device__ __forceinline__ void hash_add(unsigned long long* keys, uint32_t* vals, uint32_t cap,
                                         unsigned long long key, uint32_t w) {
    uint32_t slot = (uint32_t)mix64(key) & (cap - 1);
    while (true) {
        const unsigned long long prev = atomicCAS(&keys[slot], kEmptyKey, key);
        if (prev == kEmptyKey || prev == key) {
            atomicAdd(&vals[slot], w);
            return;
        }
        slot = (slot + 1) & (cap - 1);
    }
}
__device__ __forceinline__ uint32_t hash_get(const unsigned long long* keys, const uint32_t* vals, uint32_t cap,
                                             unsigned long long key) {
    uint32_t slot = (uint32_t)mix64(key) & (cap - 1);
    while (true) {
        const unsigned long long k = keys[slot];
        if (k == key) return vals[slot];
        if (k == kEmptyKey) return 0u;
        slot = (slot + 1) & (cap - 1);
    }
}

// IntrinsicMatrix::img_to_space_coord (types.rs:432-445) with Mat3*Vec3 of
// meancov_estimation.rs:201-216: tmp = v0*m[j][0]; tmp = tmp + v1*m[j][1]; tmp = tmp + v2*m[j][2].
__device__ __forceinline__ void img_to_space(const float* Kinv, float x, float y, float z, float out[3]) {
    float r[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float t = __fmul_rn(x, Kinv[j * 3 + 0]);
        t = __fadd_rn(t, __fmul_rn(y, Kinv[j * 3 + 1]));
        t = __fadd_rn(t, __fmul_rn(1.0f, Kinv[j * 3 + 2]));
        r[j] = t;
    }
    const float c = __fdiv_rn(z, r[2]);
    out[0] = __fmul_rn(r[0], c);
    out[1] = __fmul_rn(r[1], c);
    out[2] = __fmul_rn(r[2], c);
}
// IntrinsicMatrix::space_to_img_coord (types.rs:424-428)
__device__ __forceinline__ void space_to_img(const float* K, const float p[3], float out[2]) {
    float r[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float t = __fmul_rn(p[0], K[j * 3 + 0]);
        t = __fadd_rn(t, __fmul_rn(p[1], K[j * 3 + 1]));
        t = __fadd_rn(t, __fmul_rn(p[2], K[j * 3 + 2]));
        r[j] = t;
    }
    out[0] = __fdiv_rn(r[0], r[2]);
    out[1] = __fdiv_rn(r[1], r[2]);
}

// ================================================================ K1: summed-area table
// Row pass: exclusive prefix sums of one image row into S[y+1][0..w].  One warp per row,
// 8 pixels per lane per step (one 16-byte load), warp-shuffle scan of the lane totals.
constexpr int kSatRowWarps = 8;
__global__ void __launch_bounds__(kSatRowWarps * 32) sat_rows_kernel(const uint16_t* __restrict__ depth,
                                                                      uint32_t* __restrict__ sat, uint32_t w,
                                                                      uint32_t h, uint32_t pitch) {
    const uint32_t frame = blockIdx.y;
    const uint32_t y = blockIdx.x * kSatRowWarps + (threadIdx.x >> 5);
    if (y >= h) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint16_t* row = depth + ((size_t)frame * h + y) * w;
    uint32_t* out = sat + ((size_t)frame * (h + 1) + (y + 1)) * pitch;
    const bool vec_ok = ((w & 7u) == 0u) && ((reinterpret_cast<uintptr_t>(row) & 15u) == 0u);
    uint32_t carry = 0;
    for (uint32_t seg = 0; seg <= w; seg += 256) {  // positions 0..w inclusive
        const uint32_t x = seg + lane * 8;
        uint32_t v[8];
        if (vec_ok && x + 8 <= w) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(row + x));
            v[0] = q.x & 0xffffu; v[1] = q.x >> 16; v[2] = q.y & 0xffffu; v[3] = q.y >> 16;
            v[4] = q.z & 0xffffu; v[5] = q.z >> 16; v[6] = q.w & 0xffffu; v[7] = q.w >> 16;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (x + j < w) ? (uint32_t)__ldg(row + x + j) : 0u;
        }
        uint32_t e[8];
        uint32_t tot = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            e[j] = tot;
            tot += v[j];
        }
        uint32_t incl = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (uint32_t)d) incl += n;
        }
        const uint32_t base = carry + incl - tot;
        if (x + 7 <= w) {
            uint4 a = make_uint4(base + e[0], base + e[1], base + e[2], base + e[3]);
            uint4 b = make_uint4(base + e[4], base + e[5], base + e[6], base + e[7]);
            *reinterpret_cast<uint4*>(out + x) = a;
            *reinterpret_cast<uint4*>(out + x + 4) = b;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (x + j <= w) out[x + j] = base + e[j];
        }
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
}

// Column pass: in-place inclusive scan down every column (rows 1..h; row 0 stays zero).
// One thread per column, 8 independent loads in flight per thread.
__global__ void __launch_bounds__(128) sat_cols_kernel(uint32_t* __restrict__ sat, uint32_t w, uint32_t h,
                                                        uint32_t pitch) {
    const uint32_t frame = blockIdx.y;
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > w) return;
    uint32_t* col = sat + (size_t)frame * (h + 1) * pitch + pitch + c;  // row 1
    uint32_t acc = 0;
    uint32_t y = 0;
    for (; y + 8 <= h; y += 8) {
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = col[(size_t)(y + j) * pitch];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            acc += v[j];
            col[(size_t)(y + j) * pitch] = acc;
        }
    }
    for (; y < h; ++y) {
        acc += col[(size_t)y * pitch];
        col[(size_t)y * pitch] = acc;
    }
}

// ================================================================ K2: forest traversal
// One CTA per (tile of patches, frame).  The tile's SAT window (tw x th u32) is brought into
// shared memory by ONE TMA bulk-tensor load; all 8 taps of every node test are then shared-memory
// gathers.  Non-background patches of the tile are compacted (ballot) so that every lane of the
// traversal loop owns a live patch x tree pair; lanes of a warp are neighbouring patches of the
// same tree, so the upper levels read the same node record (broadcast) and nearby taps.
template <int kThreads>
__global__ void __launch_bounds__(kThreads) traverse_kernel(const __grid_constant__ CUtensorMap sat_map,
                                                            const NodeRec* __restrict__ nodes,
                                                            const int32_t* __restrict__ roots, int32_t* __restrict__ leaf,
                                                            FrameState* __restrict__ fs, Geometry g, TilePlan tp) {
    extern __shared__ uint8_t smem_raw[];
    // 128-byte aligned tile (TMA destination), then the barrier, then the compacted patch list
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint32_t* tile = reinterpret_cast<uint32_t*>(base);
    const uint32_t tile_bytes = tp.tw * tp.th * 4u;
    uint64_t* bar = reinterpret_cast<uint64_t*>(base + ((tile_bytes + 15u) & ~15u));
    uint16_t* live = reinterpret_cast<uint16_t*>(bar + 2);
    __shared__ uint32_t s_nlive;

    const uint32_t frame = blockIdx.y;
    const uint32_t tile_x = blockIdx.x % tp.tiles_x, tile_y = blockIdx.x / tp.tiles_x;
    const uint32_t px0 = tile_x * tp.tpx, py0 = tile_y * tp.tpy;  // first patch of the tile
    const uint32_t tid = threadIdx.x;

    if (tid == 0) {
        prefetch_tensormap(&sat_map);
        mbar_init(bar, 1);
        fence_mbar_init();
        s_nlive = 0;
    }
    __syncthreads();
    // TMA needs the innermost coordinate 16-byte aligned: start the tile at a multiple of 4
    // elements and shift the patch origins by the slack dx (the planner widened tw for it).
    const uint32_t x0 = px0 * g.stride, ax0 = x0 & ~3u, dx = x0 - ax0;
    if (tid == 0) {
        mbar_arrive_expect_tx(bar, tile_bytes);
        tma_load_3d(tile, &sat_map, (int)ax0, (int)(py0 * g.stride), (int)frame, bar);
    }
    mbar_wait(bar, 0);

    // ---- background test (prediction.rs:567-571: mean over the whole patch > 0  <=>  sum != 0)
    const uint32_t npt = tp.tpx * tp.tpy;
    const int T = (int)g.n_trees;
    int32_t* leaf_f = leaf + (size_t)frame * T * g.P;
    uint32_t my_valid = 0;
    for (uint32_t lp = tid; lp < ((npt + 31u) & ~31u); lp += kThreads) {
        bool ok = false;
        uint32_t gp = 0;
        if (lp < npt) {
            const uint32_t lx = lp % tp.tpx, ly = lp / tp.tpx;
            const uint32_t gx = px0 + lx, gy = py0 + ly;
            if (gx < g.npx && gy < g.npy) {
                gp = gy * g.npx + gx;
                const uint32_t* o = tile + ly * g.stride * tp.tw + lx * g.stride + dx;
                const uint32_t sum = o[g.sh * tp.tw + g.sw] - o[g.sw] - o[g.sh * tp.tw] + o[0];
                ok = sum != 0u;
                if (!ok)
                    for (int t = 0; t < T; ++t) leaf_f[(size_t)t * g.P + gp] = -1;
            }
        }
        const uint32_t m = __ballot_sync(0xffffffffu, ok);
        uint32_t basei = 0;
        if ((tid & 31u) == 0 && m) basei = atomicAdd(&s_nlive, (uint32_t)__popc(m));
        basei = __shfl_sync(0xffffffffu, basei, 0);
        if (ok) {
            live[basei + __popc(m & ((1u << (tid & 31u)) - 1u))] = (uint16_t)lp;
            ++my_valid;
        }
    }
    __syncthreads();
    const uint32_t nlive = s_nlive;

    // ---- root -> leaf walks: item = (tree, live patch); lanes = neighbouring live patches
    unsigned long long visits = 0;
    const uint32_t items = nlive * (uint32_t)T;
    for (uint32_t it = tid; it < items; it += kThreads) {
        const uint32_t t = it / nlive;
        const uint32_t lp = live[it - t * nlive];
        const uint32_t lx = lp % tp.tpx, ly = lp / tp.tpx;
        const uint32_t* o = tile + ly * g.stride * tp.tw + lx * g.stride + dx;
        int32_t node = __ldg(roots + t);
        while (node >= 0) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(nodes + node));
            const int2 ch = __ldg(reinterpret_cast<const int2*>(nodes + node) + 2);
            // SubImage::average_value_in_rect (types.rs:317-339) via four SAT taps per rectangle
            const uint32_t ax0 = a.x & 0xffu, ay0 = (a.x >> 8) & 0xffu, ax1 = (a.x >> 16) & 0xffu, ay1 = a.x >> 24;
            const uint32_t bx0 = a.y & 0xffu, by0 = (a.y >> 8) & 0xffu, bx1 = (a.y >> 16) & 0xffu, by1 = a.y >> 24;
            const uint32_t s1 = o[ay1 * tp.tw + ax1] - o[ay0 * tp.tw + ax1] - o[ay1 * tp.tw + ax0] + o[ay0 * tp.tw + ax0];
            const uint32_t s2 = o[by1 * tp.tw + bx1] - o[by0 * tp.tw + bx1] - o[by1 * tp.tw + bx0] + o[by0 * tp.tw + bx0];
            const uint32_t c1 = (ax1 - ax0) * (ay1 - ay0), c2 = (bx1 - bx0) * (by1 - by0);
            const double avg1 = c1 ? __ddiv_rn(__uint2double_rn(s1), __uint2double_rn(c1)) : 0.0;
            const double avg2 = c2 ? __ddiv_rn(__uint2double_rn(s2), __uint2double_rn(c2)) : 0.0;
            const double thr = __hiloint2double((int)a.w, (int)a.z);
            // HoughTreeFunctions::binarize (houghforest.rs:185-193)
            node = (__dsub_rn(avg1, avg2) > thr) ? ch.y : ch.x;
            ++visits;
        }
        const uint32_t gp = (py0 + ly) * g.npx + (px0 + lx);
        leaf_f[(size_t)t * g.P + gp] = ~node;
    }

    // ---- per-frame counters (measured mean depth feeds the roofline arithmetic)
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        visits += __shfl_xor_sync(0xffffffffu, visits, d);
        my_valid += __shfl_xor_sync(0xffffffffu, my_valid, d);
    }
    if ((tid & 31u) == 0) {
        if (visits) atomicAdd(&fs[frame].node_visits, visits);
        if (my_valid) atomicAdd(&fs[frame].n_valid, my_valid);
    }
}

// ================================================================ K3a: patch gate + hit list
__global__ void __launch_bounds__(256) gate_kernel(FrameBuffers b, Geometry g, const double* __restrict__ leaf_prob,
                                                   const LeafInfo* __restrict__ leaf_info) {
    const uint32_t frame = blockIdx.y;
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const int T = (int)g.n_trees;
    const int32_t* leaf_f = b.leaf + (size_t)frame * T * g.P;
    FrameState* fs = b.fs + frame;
    const uint32_t lane = threadIdx.x & 31u;

    bool gate = false;
    uint32_t cnt = 0;
    unsigned long long nmid = 0, nrot = 0;
    float p3[3] = {0.f, 0.f, 0.f};
    if (p < g.P) {
        const bool valid = leaf_f[p] >= 0;
        if (valid) {
            // prob = sum(leaf.prob) / len: f64 fold from 0.0 in tree order (prediction.rs:582)
            double s = 0.0;
            for (int t = 0; t < T; ++t) s = __dadd_rn(s, __ldg(leaf_prob + leaf_f[(size_t)t * g.P + p]));
            gate = __ddiv_rn(s, (double)T) > 0.7;  // prediction.rs:584
        }
        if (gate) {
            const uint32_t gx = p % g.npx, gy = p / g.npx;
            const uint32_t x = g.left_w + gx * g.stride, y = g.left_h + gy * g.stride;
            const uint16_t z = b.depth[((size_t)frame * g.h + y) * g.w + x];  // prediction.rs:551
            img_to_space(g.Kinv, (float)x, (float)y, (float)z, p3);           // prediction.rs:554
            float* o = b.p3 + ((size_t)frame * g.P + p) * 3;
            o[0] = p3[0]; o[1] = p3[1]; o[2] = p3[2];
            for (int t = 0; t < T; ++t) {
                const int32_t L = leaf_f[(size_t)t * g.P + p];
                if (!(__ldg(leaf_prob + L) > 0.0)) continue;  // prediction.rs:590
                const LeafInfo li = leaf_info[L];
                if (li.valtoadd == 0u || (li.flags & (kLeafRotOk | kLeafOffOk)) == 0u) continue;
                ++cnt;
                if (li.flags & kLeafOffOk) nmid += li.n_votes;
                if (li.flags & kLeafRotOk) nrot += li.n_votes;
            }
        }
        if (b.gate) b.gate[(size_t)frame * g.P + p] = gate ? 1 : 0;
    }
    // warp-aggregated append to the frame's hit list
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (uint32_t)d) incl += n;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    const uint32_t ngate = __popc(__ballot_sync(0xffffffffu, gate));
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        nmid += __shfl_xor_sync(0xffffffffu, nmid, d);
        nrot += __shfl_xor_sync(0xffffffffu, nrot, d);
    }
    uint32_t basei = 0;
    if (lane == 0) {
        if (total) basei = atomicAdd(&fs->n_hits, total);
        if (ngate) atomicAdd(&fs->n_gate, ngate);
        if (nmid) atomicAdd(&fs->n_mid_votes, nmid);
        if (nrot) atomicAdd(&fs->n_rot_votes, nrot);
    }
    basei = __shfl_sync(0xffffffffu, basei, 0);
    if (cnt) {
        Hit* out = b.hits + (size_t)frame * g.P * T + basei + (incl - cnt);
        for (int t = 0; t < T; ++t) {
            const int32_t L = leaf_f[(size_t)t * g.P + p];
            if (!(__ldg(leaf_prob + L) > 0.0)) continue;
            const LeafInfo li = leaf_info[L];
            if (li.valtoadd == 0u || (li.flags & (kLeafRotOk | kLeafOffOk)) == 0u) continue;
            *out++ = Hit{p, (uint32_t)L};
        }
    }
}

// ================================================================ K3b: coarse grids + seeds
struct Best {
    uint32_t val, idx;
};
__device__ __forceinline__ Best better(Best a, Best b) {
    // first maximum in scan order == strict `>` fold (prediction.rs:694-702, 733-742)
    return (b.val > a.val || (b.val == a.val && b.idx < a.idx)) ? b : a;
}
__device__ Best block_argmax(const uint32_t* cells, int n, Best* s_red) {
    Best best{0u, 0u};
    for (int i = threadIdx.x; i < n; i += blockDim.x) best = better(best, Best{cells[i], (uint32_t)i});
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        Best o{__shfl_xor_sync(0xffffffffu, best.val, d), __shfl_xor_sync(0xffffffffu, best.idx, d)};
        best = better(best, o);
    }
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        Best r = s_red[0];
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = better(r, s_red[i]);
        s_red[0] = r;
    }
    __syncthreads();
    Best r = s_red[0];
    __syncthreads();
    return r;
}

template <int G>
__global__ void __launch_bounds__(256) coarse_vote_kernel(FrameBuffers b, Geometry g, ForestDev f) {
    __shared__ uint32_t s_pos[kPosGridCells];
    __shared__ uint32_t s_rot[kRotGridCells];
    __shared__ Best s_red[8];
    __shared__ unsigned long long s_sum[8];
    __shared__ uint32_t s_cnt[8];
    __shared__ uint32_t s_last;
    const uint32_t frame = blockIdx.y;
    FrameState* fs = b.fs + frame;
    for (int i = threadIdx.x; i < kPosGridCells; i += blockDim.x) s_pos[i] = 0;
    for (int i = threadIdx.x; i < kRotGridCells; i += blockDim.x) s_rot[i] = 0;
    __syncthreads();

    const uint32_t n_hits = fs->n_hits;
    const Hit* hits = b.hits + (size_t)frame * g.P * g.n_trees;
    const float* p3f = b.p3 + (size_t)frame * g.P * 3;
    const uint32_t groups_per_block = blockDim.x / G;
    const uint32_t sub = threadIdx.x % G;
    for (uint32_t hi = blockIdx.x * groups_per_block + threadIdx.x / G; hi < n_hits; hi += gridDim.x * groups_per_block) {
        const Hit hit = hits[hi];
        const LeafInfo li = f.leaf_info[hit.leaf];
        const float px = p3f[hit.patch * 3 + 0], py = p3f[hit.patch * 3 + 1], pz = p3f[hit.patch * 3 + 2];
        for (uint32_t k = sub; k < li.n_votes; k += G) {
            const uint32_t v = li.vote_start + k;
            if (li.flags & kLeafRotOk) {  // prediction.rs:629-636
                const uint32_t bins = __ldg(f.rot_bins + v);
                const uint32_t r1 = bins & 0xffu, r2 = (bins >> 8) & 0xffu, r3 = (bins >> 16) & 0xffu;
                const uint32_t q1 = r1 * kGuessGridParts / kRotGridParts, q2 = r2 * kGuessGridParts / kRotGridParts,
                               q3 = r3 * kGuessGridParts / kRotGridParts;
                atomicAdd(&s_rot[q3 * 400 + q2 * 20 + q1], li.valtoadd);
            }
            if (li.flags & kLeafOffOk) {  // prediction.rs:644-677
                float np[3];
                np[0] = __fsub_rn(px, __ldg(f.offsets + (size_t)v * 3 + 0));
                np[1] = __fsub_rn(py, __ldg(f.offsets + (size_t)v * 3 + 1));
                np[2] = __fsub_rn(pz, __ldg(f.offsets + (size_t)v * 3 + 2));
                if (np[2] < 0.0f) continue;
                float p2[2];
                space_to_img(g.K, np, p2);
                // max!/min! macros (prediction.rs:19-25): plain comparisons, NaN -> 0.0
                const float mx = (p2[0] > 0.0f) ? p2[0] : 0.0f;
                const float x2d = (mx < (float)(g.w - 1)) ? mx : (float)(g.w - 1);
                const float my = (p2[1] > 0.0f) ? p2[1] : 0.0f;
                const float y2d = (my < (float)(g.h - 1)) ? my : (float)(g.h - 1);
                const uint32_t cx = __float2uint_rz(x2d) * kGuessGridParts / g.w;
                const uint32_t cy = __float2uint_rz(y2d) * kGuessGridParts / g.h;
                atomicAdd(&s_pos[cy * kGuessGridParts + cx], li.valtoadd);
            }
        }
    }
    __syncthreads();
    uint32_t* gpos = b.grids + (size_t)frame * (kPosGridCells + kRotGridCells);
    uint32_t* grot = gpos + kPosGridCells;
    for (int i = threadIdx.x; i < kPosGridCells; i += blockDim.x)
        if (s_pos[i]) atomicAdd(&gpos[i], s_pos[i]);
    for (int i = threadIdx.x; i < kRotGridCells; i += blockDim.x)
        if (s_rot[i]) atomicAdd(&grot[i], s_rot[i]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&fs->ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    // ---- last block of the frame: arg-max seeds (prediction.rs:694-752, 437-460)
    for (int i = threadIdx.x; i < kPosGridCells; i += blockDim.x) s_pos[i] = __ldcg(gpos + i);
    for (int i = threadIdx.x; i < kRotGridCells; i += blockDim.x) s_rot[i] = __ldcg(grot + i);
    __syncthreads();
    const Best bp = block_argmax(s_pos, kPosGridCells, s_red);
    const Best br = block_argmax(s_rot, kRotGridCells, s_red);
    const uint32_t gpw = g.w / kGuessGridParts, gph = g.h / kGuessGridParts;  // :706-707
    const uint32_t cgx = bp.idx % kGuessGridParts, cgy = bp.idx / kGuessGridParts;
    unsigned long long zs = 0;
    uint32_t zc = 0;
    const uint16_t* img = b.depth + (size_t)frame * g.h * g.w;
    for (uint32_t i = threadIdx.x; i < gpw * gph; i += blockDim.x) {
        const uint32_t xx = gpw * cgx + i % gpw, yy = gph * cgy + i / gpw;
        const uint32_t v = img[(size_t)yy * g.w + xx];
        if (v > 0) {
            zs += v;
            zc += 1;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        zs += __shfl_xor_sync(0xffffffffu, zs, d);
        zc += __shfl_xor_sync(0xffffffffu, zc, d);
    }
    if ((threadIdx.x & 31) == 0) {
        s_sum[threadIdx.x >> 5] = zs;
        s_cnt[threadIdx.x >> 5] = zc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        zs = 0;
        zc = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
            zs += s_sum[i];
            zc += s_cnt[i];
        }
        const float meanz = zc > 0 ? (float)__ddiv_rn((double)zs, (double)zc) : 0.0f;  // :721-725
        const float mxf = __fmul_rn(__fadd_rn((float)cgx, 0.5f), (float)gpw);
        const float myf = __fmul_rn(__fadd_rn((float)cgy, 0.5f), (float)gph);
        float m3[3];
        img_to_space(g.Kinv, mxf, myf, meanz, m3);
        int32_t sm[3] = {__float2int_rz(m3[0]), __float2int_rz(m3[1]), __float2int_rz(m3[2])};  // :750
        if (fs->has_guess & 1u)  // :437-441
            for (int k = 0; k < 3; ++k) sm[k] = __float2int_rz(fs->midp_guess[k]);
        const uint32_t rc[3] = {br.idx % 20u, (br.idx % 400u) / 20u, br.idx / 400u};
        for (int k = 0; k < 3; ++k) {
            // :745-747 then :458-460
            double deg = __ddiv_rn(__dadd_rn(__dmul_rn((double)rc[k], 360.0), 180.0), (double)kGuessGridParts);
            if (fs->has_guess & 2u)  // :448-450
                deg = __dadd_rn(__ddiv_rn(__dmul_rn(fs->rot_guess[k], 180.0), 3.14159), 180.0);
            fs->seed_rot[k] = __double2int_rz(__ddiv_rn(__dmul_rn(deg, (double)kRotGridParts), 360.0));
            fs->seed_mid[k] = sm[k];
        }
    }
}

// ================================================================ K4a: hash accumulators
__device__ __forceinline__ unsigned long long pow2ceil(unsigned long long v) {
    unsigned long long c = 64;
    while (c < v) c <<= 1;
    return c;
}
__global__ void plan_hash_kernel(FrameState* fs, PoolState* pool, uint32_t n_frames, unsigned long long capacity) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long off = 0;
    for (uint32_t i = 0; i < n_frames; ++i) {
        const unsigned long long nm = fs[i].n_mid_votes, nr = fs[i].n_rot_votes;
        unsigned long long cm = nm ? pow2ceil(2 * nm) : 0;
        unsigned long long cr = nr ? pow2ceil(2 * nr) : 0;
        if (cr > (1ull << 22)) cr = 1ull << 22;  // at most 120^3 distinct rotation cells
        if (cm > (1ull << 31)) cm = 1ull << 31;
        fs[i].hash_off[0] = off;
        fs[i].hash_cap[0] = (uint32_t)cm;
        off += cm;
        fs[i].hash_off[1] = off;
        fs[i].hash_cap[1] = (uint32_t)cr;
        off += cr;
    }
    pool->total_slots = off;
    pool->capacity = capacity;
    if (off > capacity) {
        pool->overflow = 1;
        for (uint32_t i = 0; i < n_frames; ++i) fs[i].hash_cap[0] = fs[i].hash_cap[1] = 0;
    } else {
        pool->overflow = 0;
    }
}
__global__ void __launch_bounds__(256) hash_clear_kernel(unsigned long long* keys, uint32_t* vals, const PoolState* pool) {
    const unsigned long long n = pool->overflow ? 0ull : pool->total_slots;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        keys[i] = kEmptyKey;
        vals[i] = 0u;
    }
}

template <int G>
__global__ void __launch_bounds__(256) insert_kernel(FrameBuffers b, Geometry g, ForestDev f, uint32_t reach) {
    const uint32_t frame = blockIdx.y;
    const FrameState* fs = b.fs + frame;
    const uint32_t cap_mid = fs->hash_cap[0], cap_rot = fs->hash_cap[1];
    unsigned long long* kmid = b.hash_keys + fs->hash_off[0];
    uint32_t* vmid = b.hash_vals + fs->hash_off[0];
    unsigned long long* krot = b.hash_keys + fs->hash_off[1];
    uint32_t* vrot = b.hash_vals + fs->hash_off[1];
    int32_t seed_mid[3] = {fs->seed_mid[0], fs->seed_mid[1], fs->seed_mid[2]};
    int32_t seed_rot[3] = {fs->seed_rot[0], fs->seed_rot[1], fs->seed_rot[2]};
    const uint32_t n_hits = fs->n_hits;
    const Hit* hits = b.hits + (size_t)frame * g.P * g.n_trees;
    const float* p3f = b.p3 + (size_t)frame * g.P * 3;
    const uint32_t groups_per_block = blockDim.x / G;
    const uint32_t sub = threadIdx.x % G;
    for (uint32_t hi = blockIdx.x * groups_per_block + threadIdx.x / G; hi < n_hits; hi += gridDim.x * groups_per_block) {
        const Hit hit = hits[hi];
        const LeafInfo li = f.leaf_info[hit.leaf];
        const float px = p3f[hit.patch * 3 + 0], py = p3f[hit.patch * 3 + 1], pz = p3f[hit.patch * 3 + 2];
        for (uint32_t k = sub; k < li.n_votes; k += G) {
            const uint32_t v = li.vote_start + k;
            unsigned long long key;
            if ((li.flags & kLeafRotOk) && cap_rot) {  // rot[(r1,r2,r3)] += valtoadd, prediction.rs:635
                const uint32_t bins = __ldg(f.rot_bins + v);
                if (make_key((int)(bins & 0xffu), (int)((bins >> 8) & 0xffu), (int)((bins >> 16) & 0xffu), seed_rot,
                             reach, &key))
                    hash_add(krot, vrot, cap_rot, key, li.valtoadd);
            }
            if ((li.flags & kLeafOffOk) && cap_mid) {  // mid[(x,y,z)] += valtoadd, prediction.rs:647-667
                const float nx = __fsub_rn(px, __ldg(f.offsets + (size_t)v * 3 + 0));
                const float ny = __fsub_rn(py, __ldg(f.offsets + (size_t)v * 3 + 1));
                const float nz = __fsub_rn(pz, __ldg(f.offsets + (size_t)v * 3 + 2));
                if (nz < 0.0f) continue;
                // z3d / ZSCALEFACTOR(=1) is exact
                if (make_key(__float2int_rz(nx), __float2int_rz(ny), __float2int_rz(nz), seed_mid, reach, &key))
                    hash_add(kmid, vmid, cap_mid, key, li.valtoadd);
            }
        }
    }
}

// ================================================================ K4b: mean-shift
// MeanShift::meanshift (meanshift.rs:328-407) for SparseArray3D<u32>: window offsets -10..+9 per
// axis, x outermost / z innermost, f32 numerators and denominator accumulated SEQUENTIALLY in that
// order (only non-zero cells contribute; adding them in order is all that matters), position
// truncated every iteration, exactly `iterations` rounds unless the denominator is exactly 0.
// A round that leaves the position unchanged makes every later round identical, so the loop
// stops there (result-neutral).
constexpr int kMsThreads = 256;
__global__ void __launch_bounds__(kMsThreads) meanshift_kernel(FrameBuffers b, const float* __restrict__ kern,
                                                               uint32_t iterations, uint32_t reach) {
    __shared__ uint32_t s_f[kKernelCells];         // cell factor, overwritten by its f32 weight
    __shared__ uint32_t s_mask[kKernelCells / 32]; // non-zero cells, bit j of word c = ord c*32+j
    __shared__ int32_t s_pos[3];
    __shared__ uint32_t s_flags, s_done;
    const uint32_t frame = blockIdx.y, which = blockIdx.x;
    FrameState* fs = b.fs + frame;
    const uint32_t cap = fs->hash_cap[which];
    const unsigned long long* keys = b.hash_keys + fs->hash_off[which];
    const uint32_t* vals = b.hash_vals + fs->hash_off[which];
    const int32_t* seedp = which == 0 ? fs->seed_mid : fs->seed_rot;
    const int32_t seed[3] = {seedp[0], seedp[1], seedp[2]};
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    constexpr int kChunks = kKernelCells / 32;  // 250
    if (tid == 0) {
        s_pos[0] = seed[0]; s_pos[1] = seed[1]; s_pos[2] = seed[2];
        s_flags = 0; s_done = 0;
    }
    __syncthreads();
    uint32_t it = 0;
    for (; it < iterations; ++it) {
        const int32_t pos[3] = {s_pos[0], s_pos[1], s_pos[2]};
        // 1. gather the 20^3 window; ord enumerates the cells in the reference loop order
        //    (x outermost, z innermost).  Each warp owns whole 32-cell chunks so that the ballot
        //    below yields the chunk's non-zero mask; non-zero cells are replaced by their weight
        //    kernel[(x+10, y+10, z+10)] * (factor as f32)  (meanshift.rs:370-377; dense kernel
        //    index z*400 + y*20 + x, meanshift.rs:78-88).
        for (uint32_t c = warp; c < (uint32_t)kChunks; c += kMsThreads / 32) {
            const uint32_t ord = c * 32 + lane;
            const uint32_t xo = ord / 400u, yo = (ord / 20u) % 20u, zo = ord % 20u;
            const int ax = (int)((uint32_t)pos[0] + (uint32_t)((int)xo - 10)),
                      ay = (int)((uint32_t)pos[1] + (uint32_t)((int)yo - 10)),
                      az = (int)((uint32_t)pos[2] + (uint32_t)((int)zo - 10));
            uint32_t fct = 0;
            unsigned long long key;
            if (make_key(ax, ay, az, seed, reach, &key)) {
                if (cap) fct = hash_get(keys, vals, cap, key);
            } else {
                atomicOr(&s_flags, 2u);  // probe outside the stored reach (never expected)
            }
            const uint32_t m = __ballot_sync(0xffffffffu, fct != 0u);
            if (lane == 0) s_mask[c] = m;
            if (fct) s_f[ord] = __float_as_uint(__fmul_rn(__ldg(kern + zo * 400u + yo * 20u + xo), (float)fct));
        }
        __syncthreads();
        // 2. sequential f32 accumulation: lanes 0..2 own the numerator components, lane 3 the
        //    denominator; each adds the non-zero cells in reference order.
        if (warp == 0) {
            float acc = 0.0f;
            if (lane < 4) {
                for (uint32_t c = 0; c < (uint32_t)kChunks; ++c) {
                    uint32_t m = s_mask[c];
                    while (m) {
                        const uint32_t ord = c * 32 + (uint32_t)(__ffs((int)m) - 1);
                        m &= m - 1;
                        const float wgt = __uint_as_float(s_f[ord]);
                        float comp = 1.0f;
                        if (lane == 0) comp = (float)(int)((uint32_t)pos[0] + (uint32_t)((int)(ord / 400u) - 10));
                        else if (lane == 1) comp = (float)(int)((uint32_t)pos[1] + (uint32_t)((int)((ord / 20u) % 20u) - 10));
                        else if (lane == 2) comp = (float)(int)((uint32_t)pos[2] + (uint32_t)((int)(ord % 20u) - 10));
                        acc = __fadd_rn(acc, __fmul_rn(comp, wgt));  // num += abs_pos * w ; den += w
                    }
                }
            }
            const float den = __shfl_sync(0xffffffffu, acc, 3);
            if (den == 0.0f) {  // "Breaking meanshift - zero sum" (meanshift.rs:385-388)
                if (lane == 0) {
                    atomicOr(&s_flags, 1u);
                    s_done = 1;
                }
            } else {
                const int np = __float2int_rz(__fdiv_rn(acc, den));  // meanshift.rs:391-394
                const int nx = __shfl_sync(0xffffffffu, np, 0), ny = __shfl_sync(0xffffffffu, np, 1),
                          nz = __shfl_sync(0xffffffffu, np, 2);
                if (lane == 0) {
                    if (b.ms_trace && it < b.ms_trace_cap) {
                        int32_t* tr = b.ms_trace + (((size_t)frame * 2 + which) * b.ms_trace_cap + it) * 3;
                        tr[0] = nx; tr[1] = ny; tr[2] = nz;
                    }
                    if (nx == pos[0] && ny == pos[1] && nz == pos[2]) s_done = 2;  // fixed point
                    s_pos[0] = nx; s_pos[1] = ny; s_pos[2] = nz;
                }
            }
        }
        __syncthreads();
        if (s_done) {
            if (s_done == 2) ++it;  // this round was executed
            break;
        }
    }
    if (tid == 0) {
        fs->ms_iters[which] = it;
        fs->ms_flags[which] = s_flags;
        dh_result* r = b.results + frame;
        if (which == 0) {  // prediction.rs:486-488
            r->mid_point[0] = (float)s_pos[0];
            r->mid_point[1] = (float)s_pos[1];
            r->mid_point[2] = (float)s_pos[2];
            r->_pad = 0;
            r->bounding_box[0] = r->bounding_box[1] = r->bounding_box[2] = r->bounding_box[3] = 0;  // :491
        } else {           // prediction.rs:477-482
            for (int k = 0; k < 3; ++k)
                r->rotation[k] = __dmul_rn(__ddiv_rn(__dsub_rn((double)s_pos[k], (double)kRotGridParts / 2.0),
                                                     (double)(kRotGridParts / 2)),
                                           3.14159);
        }
    }
}

// ================================================================ K5: per-leaf static gates
// estimate_mean_cov (meancov_estimation.rs:359-378) in reference order, one thread per leaf.
// Only the diagonal of the covariance is needed for the trace (entries are independent).
__global__ void __launch_bounds__(128) leaf_gate_kernel(const double* __restrict__ leaf_prob,
                                                        const uint32_t* __restrict__ vote_start,
                                                        const uint32_t* __restrict__ n_votes,
                                                        const float* __restrict__ offsets,
                                                        const double* __restrict__ rotations,
                                                        LeafInfo* __restrict__ out, uint32_t n_leaves) {
    const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n_leaves) return;
    const uint32_t v0 = vote_start[l], n = n_votes[l];
    LeafInfo li{v0, n, 0u, 0u};
    if (n > 0) {
        // valtoadd = ((1000.0 * prob) as usize / offsets.len()) as u32   (prediction.rs:594-595)
        const unsigned long long q = __double2ull_rz(__dmul_rn(1000.0, leaf_prob[l]));
        li.valtoadd = (uint32_t)(q / (unsigned long long)n);
        // rotations, f64
        {
            const double* s = rotations + (size_t)v0 * 3;
            double mean[3] = {s[0], s[1], s[2]};
            for (uint32_t i = 1; i < n; ++i)
                for (int k = 0; k < 3; ++k) mean[k] = __dadd_rn(mean[k], s[(size_t)i * 3 + k]);
            const double dn = (double)n;
            for (int k = 0; k < 3; ++k) mean[k] = __ddiv_rn(mean[k], dn);
            double cov[3];
            for (int k = 0; k < 3; ++k) {
                const double d = __dsub_rn(s[k], mean[k]);
                cov[k] = __dmul_rn(d, d);
            }
            for (uint32_t i = 1; i < n; ++i)
                for (int k = 0; k < 3; ++k) {
                    const double d = __dsub_rn(s[(size_t)i * 3 + k], mean[k]);
                    cov[k] = __dadd_rn(cov[k], __dmul_rn(d, d));
                }
            const double dn1 = (double)(n - 1);
            double tr = 0.0;  // trace(): fold from 0.0 (meancov_estimation.rs:260-265)
            for (int k = 0; k < 3; ++k) tr = __dadd_rn(tr, __ddiv_rn(cov[k], dn1));
            if (tr <= kMaxVarianceRot) li.flags |= kLeafRotOk;  // NaN (n == 1) -> false
        }
        // offsets, f32 with the divisors cast f64 -> f32 (meancov_estimation.rs:290-304)
        {
            const float* s = offsets + (size_t)v0 * 3;
            float mean[3] = {s[0], s[1], s[2]};
            for (uint32_t i = 1; i < n; ++i)
                for (int k = 0; k < 3; ++k) mean[k] = __fadd_rn(mean[k], s[(size_t)i * 3 + k]);
            const float dn = (float)(double)n;
            for (int k = 0; k < 3; ++k) mean[k] = __fdiv_rn(mean[k], dn);
            float cov[3];
            for (int k = 0; k < 3; ++k) {
                const float d = __fsub_rn(s[k], mean[k]);
                cov[k] = __fmul_rn(d, d);
            }
            for (uint32_t i = 1; i < n; ++i)
                for (int k = 0; k < 3; ++k) {
                    const float d = __fsub_rn(s[(size_t)i * 3 + k], mean[k]);
                    cov[k] = __fadd_rn(cov[k], __fmul_rn(d, d));
                }
            const float dn1 = (float)(double)(n - 1);
            float tr = 0.0f;
            for (int k = 0; k < 3; ++k) tr = __fadd_rn(tr, __fdiv_rn(cov[k], dn1));
            if (tr <= kMaxVarianceOffset) li.flags |= kLeafOffOk;
        }
    }
    out[l] = li;
}

// ================================================================ next-row back-ends on the same front-end
// predict_mask (prediction.rs:850-905): per non-background patch, mean prob -> u8, splat to a
// stepwidth^2 block.  Patches are visited in raster order by the reference and later patches
// overwrite earlier ones; blocks of distinct patches never overlap (block origin = centre - s/2,
// size s, centres s apart), so the order is irrelevant.
__global__ void __launch_bounds__(256) mask_kernel(FrameBuffers b, Geometry g, const double* __restrict__ leaf_prob,
                                                   uint8_t* __restrict__ mask) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.P) return;
    const int T = (int)g.n_trees;
    if (b.leaf[p] < 0) return;
    double s = 0.0;
    for (int t = 0; t < T; ++t) s = __dadd_rn(s, __ldg(leaf_prob + b.leaf[(size_t)t * g.P + p]));
    const double prob = __ddiv_rn(s, (double)T);
    const double scaled = __dmul_rn(prob, 255.0);
    // `as u8`: saturating, NaN -> 0
    const uint8_t pv = !(scaled > 0.0) ? 0 : (scaled >= 255.0 ? 255 : (uint8_t)__double2uint_rz(scaled));
    const uint32_t gx = p % g.npx, gy = p / g.npx;
    const uint32_t x = g.left_w + gx * g.stride, y = g.left_h + gy * g.stride;
    const uint32_t sw = g.stride, half = sw / 2;
    for (uint32_t j = 0; j < sw; ++j)
        for (uint32_t i = 0; i < sw; ++i) {
            if (x + i < half || y + j < half) continue;
            if (x + i - half >= g.w || y + j - half >= g.h) continue;
            mask[(size_t)(y + j - half) * g.w + (x + i - half)] = pv;
        }
}

// build_hough_image before the blur (prediction.rs:760-841).  u16 wrap-around adds are done as
// u32 atomics and truncated afterwards (sum mod 2^16 is the same either way).
__global__ void __launch_bounds__(256) hough_image_kernel(FrameBuffers b, Geometry g, ForestDev f, uint32_t* __restrict__ acc) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t T = g.n_trees;
    if (idx >= g.P * T) return;
    const uint32_t p = idx % g.P, t = idx / g.P;
    const int32_t L = b.leaf[(size_t)t * g.P + p];
    if (L < 0) return;
    const double lp = f.leaf_prob[L];
    if (!(lp >= 0.95)) return;                                             // :805
    const uint32_t v0 = f.leaf_info[L].vote_start, n = f.leaf_info[L].n_votes;
    if (n == 0) return;
    const uint32_t valtoadd = (uint32_t)(uint16_t)(__double2ull_rz(__dmul_rn(255.0, lp)) / n);  // :807-808
    const uint32_t gx = p % g.npx, gy = p / g.npx;
    const uint32_t x = g.left_w + gx * g.stride, y = g.left_h + gy * g.stride;
    float p3[3];
    img_to_space(g.Kinv, (float)x, (float)y, (float)b.depth[(size_t)y * g.w + x], p3);
    for (uint32_t k = 0; k < n; ++k) {
        float np[3], p2[2];
        for (int c = 0; c < 3; ++c) np[c] = __fsub_rn(p3[c], f.offsets[(size_t)(v0 + k) * 3 + c]);
        space_to_img(g.K, np, p2);
        const int nx = __float2int_rz(p2[0]), ny = __float2int_rz(p2[1]);  // :816
        if (nx < 0 || (uint32_t)nx >= g.w || ny < 0 || (uint32_t)ny >= g.h) continue;
        atomicAdd(&acc[(size_t)ny * g.w + nx], valtoadd);
    }
}
__global__ void narrow_u16_kernel(const uint32_t* __restrict__ in, uint16_t* __restrict__ out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint16_t)(in[i] & 0xffffu);
}

// ================================================================ debug: dump one hash table
__global__ void hash_dump_kernel(FrameBuffers b, uint32_t frame, int which, int32_t* keys_out, uint32_t* vals_out,
                                 unsigned long long* count) {
    const FrameState* fs = b.fs + frame;
    const uint32_t cap = fs->hash_cap[which];
    const unsigned long long* keys = b.hash_keys + fs->hash_off[which];
    const uint32_t* vals = b.hash_vals + fs->hash_off[which];
    const int32_t* seed = which == 0 ? fs->seed_mid : fs->seed_rot;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
        const unsigned long long k = keys[i];
        if (k == kEmptyKey) continue;
        const unsigned long long o = atomicAdd(count, 1ull);
        if (keys_out) {
            keys_out[o * 3 + 0] = (int32_t)((long long)(k & 0x1fffffull) - kKeyBias + seed[0]);
            keys_out[o * 3 + 1] = (int32_t)((long long)((k >> 21) & 0x1fffffull) - kKeyBias + seed[1]);
            keys_out[o * 3 + 2] = (int32_t)((long long)((k >> 42) & 0x1fffffull) - kKeyBias + seed[2]);
            vals_out[o] = vals[i];
        }
    }
}

// ================================================================ work counters of a pass
__global__ void __launch_bounds__(256) counters_kernel(const FrameState* __restrict__ fs, uint32_t n_frames,
                                                       unsigned long long* __restrict__ out, uint32_t P, uint32_t T,
                                                       const PoolState* __restrict__ pool) {
    if (pool->overflow) return;  // this pass will be redone with a larger pool: do not count it twice
    unsigned long long v[DH_N_COUNTERS];
    for (int k = 0; k < DH_N_COUNTERS; ++k) v[k] = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_frames; i += gridDim.x * blockDim.x) {
        const FrameState& f = fs[i];
        v[0] += 1;
        v[1] += P;
        v[2] += f.n_valid;
        v[3] += (unsigned long long)f.n_valid * T;
        v[4] += f.node_visits;
        v[5] += f.n_gate;
        v[6] += f.n_hits;
        v[7] += f.n_mid_votes;
        v[8] += f.n_rot_votes;
        v[10] += f.ms_iters[0] + f.ms_iters[1];
    }
    for (int k = 0; k < DH_N_COUNTERS; ++k) {
        unsigned long long x = v[k];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) x += __shfl_xor_sync(0xffffffffu, x, d);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(&out[k], x);
    }
}

constexpr int kTraverseThreads = 512;

}  // namespace dev

using namespace dev;

// ================================================================ launch wrappers
void launch_sat(const FrameBuffers& b, const Geometry& g, uint32_t n_frames, cudaStream_t s) {
    dim3 gr((g.h + kSatRowWarps - 1) / kSatRowWarps, n_frames);
    sat_rows_kernel<<<gr, kSatRowWarps * 32, 0, s>>>(b.depth, b.sat, g.w, g.h, g.sat_pitch);
    dim3 gc((g.w + 1 + 127) / 128, n_frames);
    sat_cols_kernel<<<gc, 128, 0, s>>>(b.sat, g.w, g.h, g.sat_pitch);
}

uint32_t traverse_smem_bytes(uint32_t tw, uint32_t th, uint32_t patches_per_tile) {
    const uint32_t tile_bytes = (tw * th * 4u + 15u) & ~15u;
    return 128u + tile_bytes + 16u + ((patches_per_tile + 31u) & ~31u) * 2u + 64u;
}

int traverse_kernel_attrs(int* regs, int* max_smem) {
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, traverse_kernel<kTraverseThreads>);
    if (e != cudaSuccess) return (int)e;
    if (regs) *regs = a.numRegs;
    if (max_smem) *max_smem = a.maxDynamicSharedSizeBytes;
    return 0;
}

void launch_traverse(const CUtensorMap& sat_map, const FrameBuffers& b, const Geometry& g, const TilePlan& tp,
                     const ForestDev& f, uint32_t n_frames, cudaStream_t s) {
    static int configured_smem = -1;
    if ((int)tp.smem_bytes > configured_smem) {
        cudaFuncSetAttribute(traverse_kernel<kTraverseThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)tp.smem_bytes);
        configured_smem = (int)tp.smem_bytes;
    }
    dim3 gr(tp.tiles_x * tp.tiles_y, n_frames);
    traverse_kernel<kTraverseThreads><<<gr, kTraverseThreads, tp.smem_bytes, s>>>(sat_map, f.nodes, f.roots, b.leaf, b.fs, g, tp);
}

void launch_gate(const FrameBuffers& b, const Geometry& g, const ForestDev& f, uint32_t n_frames, cudaStream_t s) {
    dim3 gr((g.P + 255) / 256, n_frames);
    gate_kernel<<<gr, 256, 0, s>>>(b, g, f.leaf_prob, f.leaf_info);
}

void launch_coarse(const FrameBuffers& b, const Geometry& g, const ForestDev& f, uint32_t n_frames, uint32_t splits,
                   uint32_t lanes_per_hit, cudaStream_t s) {
    dim3 gr(splits, n_frames);
    if (lanes_per_hit >= 32) coarse_vote_kernel<32><<<gr, 256, 0, s>>>(b, g, f);
    else if (lanes_per_hit >= 8) coarse_vote_kernel<8><<<gr, 256, 0, s>>>(b, g, f);
    else coarse_vote_kernel<1><<<gr, 256, 0, s>>>(b, g, f);
}

void launch_plan_and_clear(const FrameBuffers& b, uint32_t n_frames, unsigned long long capacity, cudaStream_t s) {
    plan_hash_kernel<<<1, 32, 0, s>>>(b.fs, b.pool, n_frames, capacity);
    hash_clear_kernel<<<148 * 8, 256, 0, s>>>(b.hash_keys, b.hash_vals, b.pool);
}

void launch_insert(const FrameBuffers& b, const Geometry& g, const ForestDev& f, uint32_t n_frames, uint32_t splits,
                   uint32_t lanes_per_hit, uint32_t reach, cudaStream_t s) {
    dim3 gr(splits, n_frames);
    if (lanes_per_hit >= 32) insert_kernel<32><<<gr, 256, 0, s>>>(b, g, f, reach);
    else if (lanes_per_hit >= 8) insert_kernel<8><<<gr, 256, 0, s>>>(b, g, f, reach);
    else insert_kernel<1><<<gr, 256, 0, s>>>(b, g, f, reach);
}

void launch_meanshift(const FrameBuffers& b, const ForestDev& f, uint32_t n_frames, uint32_t iterations,
                      uint32_t reach, cudaStream_t s) {
    dim3 gr(2, n_frames);
    meanshift_kernel<<<gr, kMsThreads, 0, s>>>(b, f.ms_kernel, iterations, reach);
}

void launch_leaf_gates(const double* leaf_prob, const uint32_t* vote_start, const uint32_t* n_votes,
                       const float* offsets, const double* rotations, LeafInfo* out, uint32_t n_leaves,
                       cudaStream_t s) {
    leaf_gate_kernel<<<(n_leaves + 127) / 128, 128, 0, s>>>(leaf_prob, vote_start, n_votes, offsets, rotations, out,
                                                           n_leaves);
}

void launch_mask(const FrameBuffers& b, const Geometry& g, const ForestDev& f, uint8_t* mask, cudaStream_t s) {
    mask_kernel<<<(g.P + 255) / 256, 256, 0, s>>>(b, g, f.leaf_prob, mask);
}

void launch_hough_image(const FrameBuffers& b, const Geometry& g, const ForestDev& f, uint32_t* acc32,
                        uint16_t* out16, cudaStream_t s) {
    const uint32_t n = g.P * g.n_trees;
    hough_image_kernel<<<(n + 255) / 256, 256, 0, s>>>(b, g, f, acc32);
    const uint32_t px = g.w * g.h;
    narrow_u16_kernel<<<(px + 255) / 256, 256, 0, s>>>(acc32, out16, px);
}

void launch_counters(const FrameBuffers& b, const Geometry& g, uint32_t n_frames, unsigned long long* out,
                     cudaStream_t s) {
    const uint32_t blocks = (n_frames + 255) / 256;
    counters_kernel<<<blocks, 256, 0, s>>>(b.fs, n_frames, out, g.P, g.n_trees, b.pool);
}

void launch_hash_dump(const FrameBuffers& b, uint32_t frame, int which, int32_t* keys, uint32_t* vals,
                      unsigned long long* count, cudaStream_t s) {
    hash_dump_kernel<<<256, 256, 0, s>>>(b, frame, which, keys, vals, count);
}

}  // namespace dh
