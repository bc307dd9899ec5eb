// peak_l1_gather.cu — SURVEY.md 8(d): the gather micro-benchmark that gives traverse_kernel a
// MEASURED roof.  It measures, on this box's B200, how many divergent shared-memory loads and
// divergent node-record fetches an SM completes per cycle — each alone and both together — in
// the launch shape of the traversal (1024-thread CTAs, two per SM, a ~100 KB tile each).
//
//   build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/peak_l1_gather tools/peak_l1_gather.cu
//   run:    tools/peak_l1_gather > gpurun_out/l1_peaks.jsonl          (one JSON line per variant)
//
// Every variant runs ONE wave of persistent CTAs (n_sms x 2); each thread follows `ilp` independent
// dependent chains for `steps` steps.  A step of a chain is what a node visit of the traversal is:
//   node   fetch one node record at an index that depends on the previous step
//          (tex uint4 16 B | tex uint2 8 B | ldg uint4 | ldg uint2 | none)
//   taps   two 4-byte shared-memory loads at patch origin + the node's two tap offsets (or none)
//   next   child[bit], bit = 2*(s1 - s2) > E
// pattern "tree": ten complete depth-15 trees in BFS order, all lanes of a warp start at the root
// of one tree and diverge as they descend (= the traversal's coherence profile: upper levels
// broadcast, lower levels fully divergent).  pattern "random": every step jumps to a uniformly
// random node (fully divergent at every step = the worst case / the pipe's divergent peak).
// Cycles come from clock64() inside the CTAs (max over CTAs), so the result is per SM clock and
// independent of the clock the GPU happens to run at; the CUDA-event time is printed beside it.
// A separate, untimed pass counts the shared-memory wavefronts the taps need (max number of
// distinct words per bank per warp-wide load), the figure ncu reports as
// l1tex__data_pipe_lsu_wavefronts_mem_shared.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e_ = (x);                                                                  \
        if (e_ != cudaSuccess) {                                                               \
            fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

constexpr int kThreads = 1024;
constexpr uint32_t kTw = 192, kTh = 134;           // box-sum tile of configs[1] (192 x 132) + slack
constexpr uint32_t kTileWords = kTw * kTh;         // 25 728 words = 100.5 KB
constexpr uint32_t kLevels = 15, kTrees = 10;
constexpr uint32_t kNodesPerTree = (1u << kLevels) - 1u;
constexpr uint32_t kNodes = kNodesPerTree * kTrees;  // 327 670 = the benchmark forest

enum NodeMode { kNone = 0, kTex16 = 1, kTex8 = 2, kLdg16 = 3, kLdg8 = 4, kLdg32 = 5 };
// kLdg32: one 32-byte record (LDG.E.256) holds a node AND its two children: a fetch serves two levels
constexpr uint32_t kPairLevels = 8, kRecsPerTree = 21845;  // (4^8 - 1) / 3 records of a 4-ary heap per tree
constexpr uint32_t kRecs = kRecsPerTree * kTrees;
struct alignas(32) Rec32 { uint32_t v[8]; };  // taps, E of the node; of child 0; of child 1; two spare words

struct Args {
    cudaTextureObject_t tex16, tex8;
    const uint4* tab16;
    const uint2* tab8;
    const Rec32* tab32;
    uint32_t steps;        // steps per chain
    uint32_t random;       // 0 tree pattern, 1 random pattern
    unsigned long long* cycles;   // [grid]
    unsigned long long* wavefronts;  // counting pass: [2] = {tap wavefronts, tap warp-loads}
    uint32_t* sink;
};

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t lds(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

template <int kNode, bool kTaps, int kIlp, bool kCount>
__global__ void __launch_bounds__(kThreads, 2) gather_kernel(Args a) {
    extern __shared__ __align__(16) uint32_t tile[];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    for (uint32_t i = tid; i < kTileWords; i += kThreads) tile[i] = mix(i * 2654435761u + blockIdx.x) >> 6;  // 26-bit "box sums"
    __syncthreads();
    const uint32_t tile_a = (uint32_t)__cvta_generic_to_shared(tile);
    // patch origin: lanes are neighbouring patches (stride 5), warps are patch rows
    const uint32_t org = tile_a + 4u * ((warp % 11u) * 5u * kTw + lane * 5u);
    uint32_t node[kIlp], level[kIlp], tree[kIlp], state[kIlp];
#pragma unroll
    for (int k = 0; k < kIlp; ++k) {
        tree[k] = (warp + 3u * (uint32_t)k + blockIdx.x) % kTrees;
        node[k] = tree[k] * (kNode == kLdg32 ? kRecsPerTree : kNodesPerTree);
        level[k] = 0;
        state[k] = mix(tid * 977u + blockIdx.x * 131071u + (uint32_t)k);
    }
    unsigned long long wf = 0, nld = 0;
    const long long t0 = clock64();
    for (uint32_t s = 0; s < a.steps; ++s) {
        uint4 U[kIlp];
#pragma unroll
        for (int k = 0; k < kIlp; ++k) {
            if (kNode == kTex16) U[k] = tex1Dfetch<uint4>(a.tex16, (int)node[k]);
            else if (kNode == kLdg16) U[k] = __ldg(a.tab16 + node[k]);
            else if (kNode == kTex8) { const uint2 v = tex1Dfetch<uint2>(a.tex8, (int)node[k]); U[k] = make_uint4(v.x, 0u, 0u, v.y); }
            else if (kNode == kLdg8) { const uint2 v = __ldg(a.tab8 + node[k]); U[k] = make_uint4(v.x, 0u, 0u, v.y); }
            else if (kNode == kNone) {
                // no node fetch: tap positions from a cheap per-thread generator (a handful of ALU instructions)
                state[k] = state[k] * 1664525u + 1013904223u;
                const uint32_t h = state[k] >> 8;
                U[k] = make_uint4((min(h & 63u, 56u) * kTw + min((h >> 6) & 63u, 56u)) | ((min((h >> 12) & 63u, 56u) * kTw + min((h >> 18) & 63u, 56u)) << 16),
                                  0u, 0u, (h & 0xffffu) - 0x8000u);
            }
        }
        if (kNode == kLdg32) {
#pragma unroll
            for (int k = 0; k < kIlp; ++k) {
                const Rec32 R = a.tab32[node[k]];
                const uint32_t s1 = lds(org + ((R.v[0] & 0xffffu) << 2)), s2 = lds(org + ((R.v[0] >> 16) << 2));
                const uint32_t b1 = ((int32_t)(s1 - s2) << 1) > (int32_t)R.v[1] ? 1u : 0u;
                const uint32_t t2 = b1 ? R.v[4] : R.v[2];
                const int32_t e2 = (int32_t)(b1 ? R.v[5] : R.v[3]);
                const uint32_t s3 = lds(org + ((t2 & 0xffffu) << 2)), s4 = lds(org + ((t2 >> 16) << 2));
                const uint32_t b2 = ((int32_t)(s3 - s4) << 1) > e2 ? 1u : 0u;
                const uint32_t slot = 2u * b1 + b2;
                if (a.random) {
                    node[k] = mix(node[k] * 4u + slot + lane * 0x9e3779b9u + s) % kRecs;
                } else {
                    const uint32_t local = node[k] - tree[k] * kRecsPerTree;
                    node[k] = tree[k] * kRecsPerTree + 4u * local + 1u + slot;
                    if (++level[k] == kPairLevels) {
                        level[k] = 0;
                        tree[k] = (tree[k] + 1u) % kTrees;
                        node[k] = tree[k] * kRecsPerTree;
                    }
                }
            }
            continue;
        }
#pragma unroll
        for (int k = 0; k < kIlp; ++k) {
            uint32_t bit;
            if (kTaps) {
                const uint32_t a1 = org + ((U[k].x & 0xffffu) << 2), a2 = org + ((U[k].x >> 16) << 2);
                const uint32_t s1 = lds(a1), s2 = lds(a2);
                bit = ((int32_t)(s1 - s2) << 1) > (int32_t)U[k].w ? 1u : 0u;
                if (kCount) {
                    const uint32_t act = __activemask();
                    for (int q = 0; q < 2; ++q) {
                        const uint32_t ad = q ? a2 : a1;
                        const uint32_t same_word = __match_any_sync(act, ad);
                        const uint32_t leaders = __ballot_sync(act, lane == (uint32_t)__ffs(same_word) - 1u);
                        const uint32_t same_bank = __match_any_sync(act, (ad >> 2) & 31u);
                        const uint32_t deg = __reduce_max_sync(act, (uint32_t)__popc(same_bank & leaders));
                        if (lane == 0) { wf += deg; nld += 1; }
                    }
                }
            } else {
                state[k] = state[k] * 1664525u + 1013904223u;
                bit = ((state[k] >> 16) ^ U[k].w ^ U[k].x) & 1u;
            }
            if (a.random) {
                // fully divergent: the next node is a hash of this one, the bit and the lane
                node[k] = mix(node[k] * 2u + bit + lane * 0x9e3779b9u + s) % kNodes;
            } else {
                const uint32_t local = node[k] - tree[k] * kNodesPerTree;
                node[k] = (kNode == kTex16 || kNode == kLdg16) ? (bit ? U[k].z : U[k].y) : tree[k] * kNodesPerTree + 2u * local + 1u + bit;
                if (++level[k] == kLevels) {  // reached a leaf: start the next tree at its root
                    level[k] = 0;
                    tree[k] = (tree[k] + 1u) % kTrees;
                    node[k] = tree[k] * kNodesPerTree;
                }
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < kIlp; ++k) acc ^= node[k];
    if (acc == 0xdeadbeefu) a.sink[0] = acc;
    if (tid == 0) a.cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (kCount && lane == 0) {
        atomicAdd(&a.wavefronts[0], wf);
        atomicAdd(&a.wavefronts[1], nld);
    }
}

// ---- data-stage peaks: pointer chases with (almost) no other instructions.
// kLds: two independent chases through the shared-memory tile (tile[i] = a random index into the
// tile: every warp-wide load hits random banks, ~3.5 wavefronts).  kTex: a chase through a node
// table small enough to stay in L1 (8 KB, random 16-byte fetches: texture wavefronts without
// misses).  Both together show whether the two pipes share the data stage.
// kSplit: even warps run only the shared-memory chases, odd warps only the texture chases, so each
// pipe gets its own independent instruction streams (in the combined variant above one loop
// iteration issues both and the slower pipe paces the other).
template <bool kLds, bool kTex, bool kSplit = false>
__global__ void __launch_bounds__(kThreads, 2) chase_kernel(Args a, uint32_t small_nodes) {
    extern __shared__ __align__(16) uint32_t tile[];
    const uint32_t tid = threadIdx.x;
    for (uint32_t i = tid; i < kTileWords; i += kThreads) tile[i] = (mix(i * 2654435761u + blockIdx.x) % kTileWords) << 2;  // byte offsets
    __syncthreads();
    const uint32_t tile_a = (uint32_t)__cvta_generic_to_shared(tile);
    uint32_t p0 = (mix(tid * 31u + blockIdx.x) % kTileWords) << 2, p1 = (mix(tid * 57u + 7u) % kTileWords) << 2;
    uint32_t q0 = mix(tid * 3u + blockIdx.x) % small_nodes, q1 = mix(tid * 5u + 11u) % small_nodes;
    const long long t0 = clock64();
    if (kSplit) {
        if ((tid >> 5) & 1u) {
            for (uint32_t s = 0; s < a.steps; ++s) {
                q0 = tex1Dfetch<uint4>(a.tex16, (int)q0).y;
                q1 = tex1Dfetch<uint4>(a.tex16, (int)q1).y;
            }
        } else {
            for (uint32_t s = 0; s < a.steps * 4u; ++s) {  // the shared-memory chase is ~4x faster per step: keep both halves busy to the end
                p0 = lds(tile_a + p0);
                p1 = lds(tile_a + p1);
            }
        }
    } else
    for (uint32_t s = 0; s < a.steps; ++s) {
        if (kLds) {
            p0 = lds(tile_a + p0);
            p1 = lds(tile_a + p1);
        }
        if (kTex) {
            q0 = tex1Dfetch<uint4>(a.tex16, (int)q0).y;
            q1 = tex1Dfetch<uint4>(a.tex16, (int)q1).y;
        }
    }
    const long long t1 = clock64();
    if ((p0 ^ p1 ^ q0 ^ q1) == 0xdeadbeefu) a.sink[0] = p0;
    if (tid == 0) a.cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <bool kLds, bool kTex, bool kSplit = false>
static void run_chase(const char* name, Args a, int n_sms, uint32_t steps, uint32_t small_nodes) {
    const size_t smem = (size_t)kTileWords * 4;
    CK(cudaFuncSetAttribute(chase_kernel<kLds, kTex, kSplit>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = n_sms * 2;
    a.steps = steps;
    std::vector<unsigned long long> cyc(grid);
    unsigned long long best = ~0ull;
    for (int rep = 0; rep < 3; ++rep) {
        chase_kernel<kLds, kTex, kSplit><<<grid, kThreads, smem>>>(a, small_nodes);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(cyc.data(), a.cycles, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost));
        const unsigned long long mx = *std::max_element(cyc.begin(), cyc.end());
        if (rep && mx < best) best = mx;
    }
    const double per_sm = 2.0 * kThreads * 2.0 * steps * (kSplit ? 0.5 : 1.0);  // loads of each kind per SM (split: half the warps each; LDS x4 below)
    printf("{\"variant\": \"%s\", \"pattern\": \"chase\", \"steps\": %u, \"cycles\": %llu, \"lds_thread_loads_per_cycle_per_sm\": %.3f, "
           "\"tex_thread_fetches_per_cycle_per_sm\": %.3f}\n",
           name, steps, best, kLds ? per_sm * (kSplit ? 4.0 : 1.0) / (double)best : 0.0, kTex ? per_sm / (double)best : 0.0);
    fflush(stdout);
}

struct Variant {
    const char* name;
    int node;
    bool taps;
    int ilp;
};

static int g_reps = 4;      // timed repetitions after one warm-up (argv[2]; 0 = a single launch, for runs under ncu)
static bool g_count = true; // the untimed wavefront-counting pass (argv[3])
template <int kNode, bool kTaps, int kIlp>
static void run_one(const char* name, Args a, int n_sms, bool random, uint32_t steps) {
    const size_t smem = (size_t)kTileWords * 4;
    CK(cudaFuncSetAttribute(gather_kernel<kNode, kTaps, kIlp, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(gather_kernel<kNode, kTaps, kIlp, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = n_sms * 2;
    a.steps = steps;
    a.random = random ? 1u : 0u;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best_ms = 1e30f;
    unsigned long long best_cyc = ~0ull;
    std::vector<unsigned long long> cyc(grid);
    for (int rep = g_reps ? 0 : 1; rep < (g_reps ? g_reps : 2); ++rep) {  // first repetition warms up
        CK(cudaEventRecord(e0));
        gather_kernel<kNode, kTaps, kIlp, false><<<grid, kThreads, smem>>>(a);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaMemcpy(cyc.data(), a.cycles, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost));
        const unsigned long long mx = *std::max_element(cyc.begin(), cyc.end());
        if (rep && mx < best_cyc) best_cyc = mx;
        if (rep && ms < best_ms) best_ms = ms;
    }
    double wf_per_load = 0.0;
    if (kTaps && g_count) {
        CK(cudaMemset(a.wavefronts, 0, 16));
        Args c = a;
        c.steps = std::min<uint32_t>(steps, 300u);
        gather_kernel<kNode, kTaps, kIlp, true><<<grid, kThreads, smem>>>(c);
        CK(cudaDeviceSynchronize());
        unsigned long long w[2];
        CK(cudaMemcpy(w, a.wavefronts, 16, cudaMemcpyDeviceToHost));
        wf_per_load = w[1] ? (double)w[0] / (double)w[1] : 0.0;
    }
    const double visits_per_sm = 2.0 * kThreads * (double)kIlp * steps * (kNode == kLdg32 ? 2.0 : 1.0);  // two CTAs per SM
    const double vpc = visits_per_sm / (double)best_cyc;
    printf("{\"variant\": \"%s\", \"pattern\": \"%s\", \"node\": %d, \"taps\": %d, \"ilp\": %d, \"steps\": %u, "
           "\"cycles\": %llu, \"event_ms\": %.4f, \"eff_sm_mhz\": %.0f, \"visits_per_cycle_per_sm\": %.4f, "
           "\"warp_visits_per_cycle_per_sm\": %.5f, \"tap_wavefronts_per_warp_load\": %.3f, "
           "\"tap_wavefronts_per_cycle_per_sm\": %.4f}\n",
           name, random ? "random" : "tree", kNode, kTaps ? 2 : 0, kIlp, steps, best_cyc, best_ms, (double)best_cyc / (best_ms * 1e3), vpc,
           vpc / 32.0, wf_per_load, kTaps ? vpc / 32.0 * 2.0 * wf_per_load : 0.0);
    fflush(stdout);
    CK(cudaEventDestroy(e0));
    CK(cudaEventDestroy(e1));
}

int main(int argc, char** argv) {
    uint32_t steps = argc > 1 ? (uint32_t)atoi(argv[1]) : 1500u;
    if (argc > 2) g_reps = atoi(argv[2]);
    if (argc > 3) g_count = atoi(argv[3]) != 0;
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    const int n_sms = prop.multiProcessorCount;
    fprintf(stderr, "%s, %d SMs, max clock %d kHz\n", prop.name, n_sms, prop.clockRate);
    // node tables: ten complete trees in BFS order; taps = two random 24x24 rectangle origins inside an
    // 80x80 patch (offset y*tw + x, x,y in [0,56]); child links explicit in the 16-byte record
    std::vector<uint4> h16(kNodes);
    std::vector<uint2> h8(kNodes);
    uint64_t rng = 0x1234567887654321ull;
    auto next = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return (uint32_t)(rng >> 16); };
    for (uint32_t t = 0; t < kTrees; ++t)
        for (uint32_t i = 0; i < kNodesPerTree; ++i) {
            const uint32_t x1 = next() % 57u, y1 = next() % 57u, x2 = next() % 57u, y2 = next() % 57u;
            const uint32_t taps = (y1 * kTw + x1) | ((y2 * kTw + x2) << 16);
            const int32_t E = (int32_t)(next() % 600000u) - 300000;
            const uint32_t g = t * kNodesPerTree + i;
            h16[g] = make_uint4(taps, t * kNodesPerTree + std::min(2u * i + 1u, kNodesPerTree - 1u),
                                t * kNodesPerTree + std::min(2u * i + 2u, kNodesPerTree - 1u), (uint32_t)E);
            h8[g] = make_uint2(taps, (uint32_t)E);
        }
    std::vector<Rec32> h32(kRecs);
    for (uint32_t g = 0; g < kRecs; ++g)
        for (int j = 0; j < 3; ++j) {
            const uint32_t x1 = next() % 57u, y1 = next() % 57u, x2 = next() % 57u, y2 = next() % 57u;
            h32[g].v[2 * j] = (y1 * kTw + x1) | ((y2 * kTw + x2) << 16);
            h32[g].v[2 * j + 1] = (uint32_t)((int32_t)(next() % 600000u) - 300000);
        }
    Args a{};
    Rec32* d32;
    CK(cudaMalloc(&d32, sizeof(Rec32) * kRecs));
    CK(cudaMemcpy(d32, h32.data(), sizeof(Rec32) * kRecs, cudaMemcpyHostToDevice));
    a.tab32 = d32;
    uint4* d16;
    uint2* d8;
    CK(cudaMalloc(&d16, sizeof(uint4) * kNodes));
    CK(cudaMalloc(&d8, sizeof(uint2) * kNodes));
    CK(cudaMemcpy(d16, h16.data(), sizeof(uint4) * kNodes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d8, h8.data(), sizeof(uint2) * kNodes, cudaMemcpyHostToDevice));
    a.tab16 = d16;
    a.tab8 = d8;
    {
        cudaResourceDesc rd{};
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = d16;
        rd.res.linear.desc = cudaCreateChannelDesc<uint4>();
        rd.res.linear.sizeInBytes = sizeof(uint4) * kNodes;
        cudaTextureDesc td{};
        td.readMode = cudaReadModeElementType;
        CK(cudaCreateTextureObject(&a.tex16, &rd, &td, nullptr));
        rd.res.linear.devPtr = d8;
        rd.res.linear.desc = cudaCreateChannelDesc<uint2>();
        rd.res.linear.sizeInBytes = sizeof(uint2) * kNodes;
        CK(cudaCreateTextureObject(&a.tex8, &rd, &td, nullptr));
    }
    CK(cudaMalloc(&a.cycles, sizeof(unsigned long long) * n_sms * 2));
    CK(cudaMalloc(&a.wavefronts, 16));
    CK(cudaMalloc(&a.sink, 4));
    {
        // the small table: nodes 0..511 (8 KB) linked among themselves at random
        std::vector<uint4> small(512);
        for (uint32_t i = 0; i < 512; ++i) small[i] = make_uint4(next(), next() % 512u, next() % 512u, next());
        CK(cudaMemcpy(d16, small.data(), sizeof(uint4) * 512, cudaMemcpyHostToDevice));
        run_chase<true, false>("chase_lds", a, n_sms, steps, 512);
        run_chase<false, true>("chase_tex16_l1", a, n_sms, steps, 512);
        run_chase<true, true>("chase_lds_tex16_l1", a, n_sms, steps, 512);
        run_chase<true, true, true>("chase_split_lds_tex16_l1", a, n_sms, steps, 512);
        CK(cudaMemcpy(d16, h16.data(), sizeof(uint4) * kNodes, cudaMemcpyHostToDevice));
    }
    if (argc > 4 && atoi(argv[4]) == 0) return 0;  // chases only
    for (int random = 0; random < 2; ++random) {
        const bool r = random != 0;
        // (a) divergent shared-memory taps alone
        run_one<kNone, true, 1>("lds_only_ilp1", a, n_sms, r, steps);
        run_one<kNone, true, 2>("lds_only_ilp2", a, n_sms, r, steps);
        run_one<kNone, true, 4>("lds_only_ilp4", a, n_sms, r, steps);
        // (b) node fetches alone
        run_one<kTex16, false, 1>("tex16_only_ilp1", a, n_sms, r, steps);
        run_one<kTex16, false, 2>("tex16_only_ilp2", a, n_sms, r, steps);
        run_one<kTex16, false, 4>("tex16_only_ilp4", a, n_sms, r, steps);
        run_one<kTex8, false, 1>("tex8_only_ilp1", a, n_sms, r, steps);
        run_one<kTex8, false, 4>("tex8_only_ilp4", a, n_sms, r, steps);
        run_one<kLdg16, false, 1>("ldg16_only_ilp1", a, n_sms, r, steps);
        run_one<kLdg16, false, 4>("ldg16_only_ilp4", a, n_sms, r, steps);
        run_one<kLdg8, false, 1>("ldg8_only_ilp1", a, n_sms, r, steps);
        run_one<kLdg8, false, 4>("ldg8_only_ilp4", a, n_sms, r, steps);
        // (d) two levels per 32-byte record (one LDG.E.256 + four taps): counted as TWO visits per step
        run_one<kLdg32, true, 1>("ldg32x2_lds_ilp1", a, n_sms, r, steps / 2);
        run_one<kLdg32, true, 2>("ldg32x2_lds_ilp2", a, n_sms, r, steps / 2);
        // (c) both: the skeleton of a node visit
        run_one<kTex16, true, 1>("tex16_lds_ilp1", a, n_sms, r, steps);
        run_one<kTex16, true, 2>("tex16_lds_ilp2", a, n_sms, r, steps);
        run_one<kTex16, true, 4>("tex16_lds_ilp4", a, n_sms, r, steps);
        run_one<kTex8, true, 1>("tex8_lds_ilp1", a, n_sms, r, steps);
        run_one<kTex8, true, 2>("tex8_lds_ilp2", a, n_sms, r, steps);
        run_one<kLdg16, true, 1>("ldg16_lds_ilp1", a, n_sms, r, steps);
        run_one<kLdg16, true, 2>("ldg16_lds_ilp2", a, n_sms, r, steps);
        run_one<kLdg8, true, 1>("ldg8_lds_ilp1", a, n_sms, r, steps);
        run_one<kLdg8, true, 2>("ldg8_lds_ilp2", a, n_sms, r, steps);
    }
    return 0;
}
