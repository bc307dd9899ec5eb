// dh_kernels.cuh — launch interfaces of the sm_100a kernels (definitions in dh_kernels.cu).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "dh_types.hpp"
#include "../../include/depthhead_cuda.h"

namespace dh {

// Per-frame device bookkeeping, zeroed at the start of every pipeline pass.
struct alignas(16) FrameState {
    uint32_t n_hits;          // voting patch*tree pairs appended to the hit list
    uint32_t n_valid;         // non-background patches
    uint32_t n_gate;          // patches with mean prob > 0.7
    uint32_t ticket;          // last-block detection in coarse_vote_kernel
    unsigned long long n_mid_votes;  // centre votes the hits will cast (upper bound of table load)
    unsigned long long n_rot_votes;
    unsigned long long node_visits;
    unsigned long long hash_off[2];  // slot offset of the centre / rotation table in the pool
    uint32_t hash_cap[2];            // slots (power of two, 0 = empty table)
    int32_t seed_mid[3];
    int32_t seed_rot[3];
    uint32_t ms_iters[2];            // mean-shift iterations actually executed
    uint32_t ms_flags[2];            // bit0 zero-sum break, bit1 probe outside the stored reach
    uint32_t has_guess;              // bit0 midp_guess, bit1 rot_guess supplied by the caller
    float midp_guess[3];
    double rot_guess[3];
    uint32_t _pad[2];
};

struct PoolState {
    unsigned long long total_slots;  // slots needed by this pass
    unsigned long long capacity;     // slots available
    uint32_t overflow;               // 1 = pool too small, tables disabled for this pass
    uint32_t _pad;
};

struct Geometry {
    uint32_t w, h;            // image
    uint32_t sw, sh;          // sub-image (patch)
    uint32_t left_w, left_h;  // sw/2, sh/2 (prediction.rs:535-538)
    uint32_t stride;          // stepwidth
    uint32_t npx, npy, P;     // patches per row / column / frame
    uint32_t sat_pitch;       // elements per SAT row (multiple of 4)
    uint32_t n_trees;
    float K[9], Kinv[9];
};

struct TilePlan {
    uint32_t tpx, tpy;        // patches per tile
    uint32_t tiles_x, tiles_y;
    uint32_t tw, th;          // SAT tile extent in elements (tw multiple of 4)
    uint32_t smem_bytes;
    uint32_t threads;
};

struct ForestDev {
    const NodeRec* nodes;
    const int32_t* roots;
    const double* leaf_prob;
    const LeafInfo* leaf_info;
    const float* offsets;      // 3 per vote
    const uint32_t* rot_bins;  // per vote
    const float* ms_kernel;    // 8000
    int32_t n_trees;
};

struct FrameBuffers {
    const uint16_t* depth;  // [F][h][w]
    uint32_t* sat;          // [F][h+1][pitch]
    int32_t* leaf;          // [F][T][P]
    float* p3;              // [F][P][3]
    uint8_t* gate;          // [F][P]
    Hit* hits;              // [F][P*T]
    uint32_t* grids;        // [F][400 + 8000]
    FrameState* fs;         // [F]
    unsigned long long* hash_keys;
    uint32_t* hash_vals;
    PoolState* pool;
    dh_result* results;     // [F]
    int32_t* ms_trace;      // [F][2][iters][3] or nullptr
    uint32_t ms_trace_cap;  // iterations per trace
};

void launch_sat(const FrameBuffers& b, const Geometry& g, uint32_t n_frames, cudaStream_t s);
void launch_traverse(const CUtensorMap& sat_map, const FrameBuffers& b, const Geometry& g, const TilePlan& tp,
                     const ForestDev& f, uint32_t n_frames, cudaStream_t s);
void launch_gate(const FrameBuffers& b, const Geometry& g, const ForestDev& f, uint32_t n_frames, cudaStream_t s);
void launch_coarse(const FrameBuffers& b, const Geometry& g, const ForestDev& f, uint32_t n_frames, uint32_t splits,
                   uint32_t lanes_per_hit, cudaStream_t s);
void launch_plan_and_clear(const FrameBuffers& b, uint32_t n_frames, unsigned long long capacity, cudaStream_t s);
void launch_insert(const FrameBuffers& b, const Geometry& g, const ForestDev& f, uint32_t n_frames, uint32_t splits,
                   uint32_t lanes_per_hit, uint32_t reach, cudaStream_t s);
void launch_meanshift(const FrameBuffers& b, const ForestDev& f, uint32_t n_frames, uint32_t iterations,
                      uint32_t reach, cudaStream_t s);
void launch_leaf_gates(const double* leaf_prob, const uint32_t* vote_start, const uint32_t* n_votes,
                       const float* offsets, const double* rotations, LeafInfo* out, uint32_t n_leaves,
                       cudaStream_t s);
void launch_mask(const FrameBuffers& b, const Geometry& g, const ForestDev& f, uint8_t* mask, cudaStream_t s);
void launch_hough_image(const FrameBuffers& b, const Geometry& g, const ForestDev& f, uint32_t* acc32,
                        uint16_t* out16, cudaStream_t s);
void launch_counters(const FrameBuffers& b, const Geometry& g, uint32_t n_frames, unsigned long long* out,
                     cudaStream_t s);
void launch_hash_dump(const FrameBuffers& b, uint32_t frame, int which, int32_t* keys, uint32_t* vals,
                      unsigned long long* count, cudaStream_t s);

// smem bytes the traversal kernel needs for a tile (tile + barrier + alignment slack)
uint32_t traverse_smem_bytes(uint32_t tw, uint32_t th, uint32_t patches_per_tile);
int traverse_kernel_attrs(int* regs, int* max_smem);

}  // namespace dh
