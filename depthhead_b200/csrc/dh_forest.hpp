// dh_forest.hpp — host-side model: HoughPrediction (prediction.rs:239-256) flattened once into
// structure-of-arrays tables ready for upload.  Immutable after load except for the three
// scalars the reference lets callers change (stepwidth, meanshift_iterations, sigma).
#pragma once

#include <atomic>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "dh_types.hpp"

namespace dh {

struct ModelError : std::runtime_error {
    int code;
    ModelError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// The loader-facing raw form: exactly what the JSON document (or dh_forest_from_arrays) holds.
// Keeping this separate from HostForest is what makes the `forest` sub-schema swappable: a
// different container layout (e.g. stamm's real one, should its source become available) only
// needs another function producing RawForest.
struct RawForest {
    uint32_t stepwidth = 0, subimage_width = 0, subimage_height = 0, meanshift_iterations = 0;
    float gaussian_sigma = 0.0f;
    int32_t n_trees = 0;
    std::vector<int64_t> tree_node_off{0}, tree_leaf_off{0};
    std::vector<int32_t> rects;     // 8 per node (file order)
    std::vector<double> threshold;  // per node
    std::vector<int32_t> child;     // 2 per node, tree-local
    std::vector<double> prob;       // per leaf (file order)
    std::vector<int64_t> vote_off{0};
    std::vector<float> offsets;     // 3 per vote
    std::vector<double> rotations;  // 3 per vote
};

RawForest parse_hough_prediction_json(const char* json, size_t len);

struct HostForest {
    // scalars (prediction.rs:240-255)
    std::atomic<uint32_t> stepwidth{0};
    std::atomic<uint32_t> meanshift_iterations{0};
    uint32_t subimage_width = 0, subimage_height = 0;
    float gaussian_sigma = 0.0f;
    std::atomic<uint64_t> sigma_version{1};  // bumped by update_sigma -> contexts rebuild the kernel table

    int32_t n_trees = 0;
    int32_t max_depth = 0;                 // longest root->leaf path in nodes (for work estimates)
    // Every feature rectangle of every node has this size (0 = sizes differ).  True for forests
    // trained by the reference: houghforest.rs:230-234 draws both rectangles with one scale factor
    // (hough_tree_trainer.rs:165 sets min = max = 0.3 -> 24x24 inside 80x80).  Enables the
    // box-sum traversal (two taps per node instead of eight).
    uint32_t uniform_rw = 0, uniform_rh = 0;
    std::vector<NodeRec> nodes;            // BFS order per tree, trees back to back
    std::vector<int32_t> roots;            // per tree: >=0 global node index, <0 ~global leaf id
    std::vector<int64_t> tree_node_off, tree_leaf_off;
    std::vector<double> leaf_prob;         // global leaf id = tree_leaf_off[t] + file index
    std::vector<uint32_t> leaf_vote_start; // per leaf
    std::vector<uint32_t> leaf_n_votes;    // per leaf
    std::vector<float> offsets;            // 3 per vote, mm
    std::vector<double> rotations;         // 3 per vote, degrees (kept for the mean/cov kernel)
    std::vector<uint32_t> rot_bins;        // per vote: r1 | r2<<8 | r3<<16, each in [0,120)
    uint32_t max_votes_per_leaf = 0;
    uint64_t serial = 0;                   // unique id: contexts key their device copy on it
    // HoughTreeFunctions (houghforest.rs:89-122) written into the document by forest_to_json; the
    // prediction path never reads them.  Set by the trainer; defaults are the reference trainer's.
    double fn_min_subrect_factor = 0.3, fn_max_subrect_factor = 0.3, fn_steepness = 5.0;
    uint64_t fn_number_of_gen_features = 2000, fn_max_depth = 15, fn_min_subset_size = 20;

    size_t n_nodes() const { return nodes.size(); }
    size_t n_leaves() const { return leaf_prob.size(); }
    size_t n_votes() const { return rot_bins.size(); }
};

// Validates (everything the reference would panic / loop forever on becomes an error) and
// re-lays the nodes out in BFS order.  Throws ModelError.
HostForest* flatten_forest(const RawForest& raw);

// serde_json::to_string(&HoughPrediction) (hough_tree_trainer.rs:182 `tojson`): the document
// parse_hough_prediction_json reads, floats in shortest round-trip form like serde_json (ryu).
std::string forest_to_json(const HostForest& hf);

// Mat3::inv (meancov_estimation.rs:344-352), f32, adjugate / det.
void mat3_inverse_f32(const float k[9], float inv[9]);
// FullArray3D::build_kernel(20, sigma) (meanshift.rs:228-252): 8000 f32, index z*400+y*20+x.
void build_meanshift_kernel(float sigma, float* out8000);
// imageproc's gaussian_kernel_f32(sigma): 2 * ceil(2 sigma) + 1 taps (build_hough_image's blur, prediction.rs:844)
std::vector<float> build_gaussian_blur_kernel(float sigma);

// Biwi side files (dh_biwi.cpp; src/db_reader/biwi.rs:27-86)
void biwi_depth_dims(const uint8_t* file, size_t len, uint32_t* w, uint32_t* h);
void biwi_parse_cal(const char* text, size_t len, float K[9]);
void biwi_parse_pose(const uint8_t* file, size_t len, const float K[9], float pos3d[3], float pos2d[2], float rot[3]);

}  // namespace dh
