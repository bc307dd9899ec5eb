// dh_forest.cpp — load + validate + flatten the HoughPrediction model (host only).
//
// Schema.  The depthhead-owned structs follow the reference's serde derives exactly:
//   HoughPrediction      prediction.rs:239-256   {stepwidth, subimage_width, subimage_height,
//                                                 gaussian_sigma, forest, meanshift_iterations}
//   NodeParam            houghforest.rs:63-68    {"r1": Rect, "r2": Rect, "threshold": f64}
//   Rect                 types.rs:33-37          {"topleft":[x,y], "bottomright":[x,y]}
//   LeafParam            houghforest.rs:73-78    {"prob": f64, "offsets": [[f32;3]..], "rotations": [[f64;3]..]}
//   HoughTreeFunctions   houghforest.rs:89-122   parsed and ignored (prediction never reads it)
// The `forest` container is stamm 0.2.0's RandomForest (Cargo.lock:1154-1162) whose serde layout
// cannot be read offline, so this build DEFINES it:
//   "forest": {"trees": [ {"functions": {..}, "nodes": [ {"param": NodeParam, "children": [c0, c1]} .. ],
//                          "leaves": [ LeafParam .. ]} .. ]}
// children[bit] is taken with bit = (avg1 - avg2 > threshold) (houghforest.rs:185-193); a child
// >= 0 is a node index inside the tree (root = node 0), a child < 0 is ~leaf index; a tree with
// no nodes is the single leaf 0.  Only parse_forest_container() knows this layout.
#include "dh_forest.hpp"

#include <charconv>
#include <cmath>
#include <cstring>
#include <deque>
#include <limits>
#include <memory>
#include <algorithm>

#include "dh_json.hpp"
#include "../../include/depthhead_cuda.h"

namespace dh {

namespace {

struct Seen {
    unsigned bits = 0;
    void mark(JsonReader& r, unsigned b, const char* name) {
        if (bits & (1u << b)) r.fail(std::string("duplicate field `") + name + "`");
        bits |= 1u << b;
    }
    void require(JsonReader& r, unsigned b, const char* name) const {
        if (!(bits & (1u << b))) r.fail(std::string("missing field `") + name + "`");
    }
};

void parse_rect(JsonReader& r, int32_t out[4]) {
    Seen seen;
    r.parse_object([&](const std::string& k) {
        auto pair = [&](int32_t* dst) {
            size_t n = 0;
            r.parse_array([&](size_t i) {
                uint32_t v = r.parse_u32();
                if (i < 2) dst[i] = (int32_t)std::min<uint32_t>(v, 0x7FFFFFFFu);
                n = i + 1;
            });
            if (n != 2) r.fail("Rect corner must have 2 elements");
        };
        if (k == "topleft") { seen.mark(r, 0, "topleft"); pair(out); }
        else if (k == "bottomright") { seen.mark(r, 1, "bottomright"); pair(out + 2); }
        else r.skip_value();
    });
    seen.require(r, 0, "topleft");
    seen.require(r, 1, "bottomright");
}

void parse_node_param(JsonReader& r, RawForest& f) {
    int32_t rc[8];
    double thr = 0.0;
    Seen seen;
    r.parse_object([&](const std::string& k) {
        if (k == "r1") { seen.mark(r, 0, "r1"); parse_rect(r, rc); }
        else if (k == "r2") { seen.mark(r, 1, "r2"); parse_rect(r, rc + 4); }
        else if (k == "threshold") { seen.mark(r, 2, "threshold"); thr = r.parse_f64(); }
        else r.skip_value();
    });
    seen.require(r, 0, "r1");
    seen.require(r, 1, "r2");
    seen.require(r, 2, "threshold");
    f.rects.insert(f.rects.end(), rc, rc + 8);
    f.threshold.push_back(thr);
}

template <typename T, typename ParseFn>
size_t parse_vec3_list(JsonReader& r, std::vector<T>& dst, ParseFn&& parse_one) {
    size_t count = 0;
    r.parse_array([&](size_t) {
        size_t n = 0;
        T v[3] = {0, 0, 0};
        r.parse_array([&](size_t i) {
            T x = parse_one();
            if (i < 3) v[i] = x;
            n = i + 1;
        });
        if (n != 3) r.fail("Vec3 must have 3 elements");
        dst.insert(dst.end(), v, v + 3);
        ++count;
    });
    return count;
}

void parse_leaf(JsonReader& r, RawForest& f) {
    Seen seen;
    double prob = 0.0;
    size_t n_off = 0, n_rot = 0;
    r.parse_object([&](const std::string& k) {
        if (k == "prob") { seen.mark(r, 0, "prob"); prob = r.parse_f64(); }
        else if (k == "offsets") { seen.mark(r, 1, "offsets"); n_off = parse_vec3_list<float>(r, f.offsets, [&] { return r.parse_f32(); }); }
        else if (k == "rotations") { seen.mark(r, 2, "rotations"); n_rot = parse_vec3_list<double>(r, f.rotations, [&] { return r.parse_f64(); }); }
        else r.skip_value();
    });
    seen.require(r, 0, "prob");
    seen.require(r, 1, "offsets");
    seen.require(r, 2, "rotations");
    if (n_off != n_rot)
        r.fail("leaf with offsets.len() != rotations.len() (never produced by comp_leaf_data, houghforest.rs:204-225)");
    f.prob.push_back(prob);
    f.vote_off.push_back(f.vote_off.back() + (int64_t)n_off);
}

// The only function that knows the builder-defined layout of the stamm container.
void parse_forest_container(JsonReader& r, RawForest& f) {
    Seen seen;
    r.parse_object([&](const std::string& k) {
        if (k != "trees") { r.skip_value(); return; }
        seen.mark(r, 0, "trees");
        r.parse_array([&](size_t) {
            Seen ts;
            int64_t nodes_before = (int64_t)f.threshold.size();
            int64_t leaves_before = (int64_t)f.prob.size();
            r.parse_object([&](const std::string& tk) {
                if (tk == "nodes") {
                    ts.mark(r, 0, "nodes");
                    r.parse_array([&](size_t) {
                        Seen ns;
                        r.parse_object([&](const std::string& nk) {
                            if (nk == "param") { ns.mark(r, 0, "param"); parse_node_param(r, f); }
                            else if (nk == "children") {
                                ns.mark(r, 1, "children");
                                size_t n = 0;
                                int64_t c[2] = {0, 0};
                                r.parse_array([&](size_t i) {
                                    int64_t v = r.parse_i64();
                                    if (v > INT32_MAX || v < INT32_MIN) r.fail("child index out of range");
                                    if (i < 2) c[i] = v;
                                    n = i + 1;
                                });
                                if (n != 2) r.fail("children must have 2 elements");
                                f.child.push_back((int32_t)c[0]);
                                f.child.push_back((int32_t)c[1]);
                            } else r.skip_value();
                        });
                        ns.require(r, 0, "param");
                        ns.require(r, 1, "children");
                    });
                } else if (tk == "leaves") {
                    ts.mark(r, 1, "leaves");
                    r.parse_array([&](size_t) { parse_leaf(r, f); });
                } else {
                    r.skip_value();  // "functions": HoughTreeFunctions — prediction ignores it
                }
            });
            ts.require(r, 0, "nodes");
            ts.require(r, 1, "leaves");
            if ((int64_t)f.threshold.size() * 2 != (int64_t)f.child.size()) r.fail("node without children");
            (void)nodes_before;
            (void)leaves_before;
            f.tree_node_off.push_back((int64_t)f.threshold.size());
            f.tree_leaf_off.push_back((int64_t)f.prob.size());
            f.n_trees++;
        });
    });
    seen.require(r, 0, "trees");
}

// Rust `as` casts (saturating, NaN -> 0)
inline int32_t rs_f64_as_i32(double v) {
    if (std::isnan(v)) return 0;
    if (v >= 2147483648.0) return INT32_MAX;
    if (v <= -2147483649.0) return INT32_MIN;
    return (int32_t)v;
}

}  // namespace

RawForest parse_hough_prediction_json(const char* json, size_t len) {
    JsonReader r(json, len);
    RawForest f;
    Seen seen;
    r.parse_object([&](const std::string& k) {
        if (k == "stepwidth") { seen.mark(r, 0, "stepwidth"); f.stepwidth = r.parse_u32(); }
        else if (k == "subimage_width") { seen.mark(r, 1, "subimage_width"); f.subimage_width = r.parse_u32(); }
        else if (k == "subimage_height") { seen.mark(r, 2, "subimage_height"); f.subimage_height = r.parse_u32(); }
        else if (k == "gaussian_sigma") { seen.mark(r, 3, "gaussian_sigma"); f.gaussian_sigma = r.parse_f32(); }
        else if (k == "forest") { seen.mark(r, 4, "forest"); parse_forest_container(r, f); }
        else if (k == "meanshift_iterations") { seen.mark(r, 5, "meanshift_iterations"); f.meanshift_iterations = r.parse_u32(); }
        else r.skip_value();  // kernel3d is skip_serializing/skip_deserializing (prediction.rs:252)
    });
    r.expect_end();
    seen.require(r, 0, "stepwidth");
    seen.require(r, 1, "subimage_width");
    seen.require(r, 2, "subimage_height");
    seen.require(r, 3, "gaussian_sigma");
    seen.require(r, 4, "forest");
    seen.require(r, 5, "meanshift_iterations");
    return f;
}

static uint64_t next_serial() {
    static std::atomic<uint64_t> s{1};
    return s.fetch_add(1);
}

HostForest* flatten_forest(const RawForest& raw) {
    auto bad = [](int code, const std::string& m) -> void { throw ModelError(code, m); };
    if (raw.n_trees < 1) bad(DH_E_JSON, "forest has no trees");
    if (raw.subimage_width < 1 || raw.subimage_height < 1 || raw.subimage_width > 255 || raw.subimage_height > 255)
        bad(DH_E_SHAPE, "subimage_width/height must be in 1..255 (rect coordinates are packed as u8; the u32 "
                        "summed-area table is exact for patches up to 255x255)");
    if (raw.stepwidth < 1) bad(DH_E_SHAPE, "stepwidth 0: the reference's sliding window never advances (prediction.rs:684)");
    if (raw.meanshift_iterations > 65535) bad(DH_E_ARG, "meanshift_iterations > 65535 is not supported");
    if (!(raw.gaussian_sigma == raw.gaussian_sigma)) bad(DH_E_JSON, "gaussian_sigma is NaN");
    const int T = raw.n_trees;
    if ((int)raw.tree_node_off.size() != T + 1 || (int)raw.tree_leaf_off.size() != T + 1)
        bad(DH_E_ARG, "tree offset arrays must have n_trees+1 entries");
    const int64_t NN = raw.tree_node_off[T], NL = raw.tree_leaf_off[T];
    if (NN < 0 || NL < 1 || NN > 0x7FFFFFF0ll || NL > 0x7FFFFFF0ll) bad(DH_E_JSON, "node/leaf count out of range");
    if ((int64_t)raw.threshold.size() != NN || (int64_t)raw.rects.size() != NN * 8 || (int64_t)raw.child.size() != NN * 2 ||
        (int64_t)raw.prob.size() != NL || (int64_t)raw.vote_off.size() != NL + 1)
        bad(DH_E_ARG, "forest array sizes are inconsistent");
    const int64_t NV = raw.vote_off[NL];
    if (NV < 0 || NV > 0xFFFFFFF0ll || (int64_t)raw.offsets.size() != NV * 3 || (int64_t)raw.rotations.size() != NV * 3)
        bad(DH_E_ARG, "vote arrays are inconsistent / too many votes");

    std::unique_ptr<HostForest> hf(new HostForest());
    hf->stepwidth = raw.stepwidth;
    hf->meanshift_iterations = raw.meanshift_iterations;
    hf->subimage_width = raw.subimage_width;
    hf->subimage_height = raw.subimage_height;
    hf->gaussian_sigma = raw.gaussian_sigma;
    hf->n_trees = T;
    hf->serial = next_serial();
    hf->tree_leaf_off = raw.tree_leaf_off;
    hf->tree_node_off.assign(1, 0);
    hf->nodes.reserve((size_t)NN);
    hf->roots.resize(T);

    // ---- nodes: validate, BFS re-layout (children of the upper levels become contiguous, so the
    // hot top of every tree shares cache lines), local -> global indices.
    std::vector<int32_t> newidx;  // tree-local old index -> global new index
    std::vector<int32_t> order;   // BFS order (old local indices)
    std::vector<int32_t> depth_of;
    int32_t max_depth = 0;
    for (int t = 0; t < T; ++t) {
        const int64_t n0 = raw.tree_node_off[t], n1 = raw.tree_node_off[t + 1];
        const int64_t l0 = raw.tree_leaf_off[t], l1 = raw.tree_leaf_off[t + 1];
        if (n1 < n0 || l1 <= l0) bad(DH_E_JSON, "tree " + std::to_string(t) + " has no leaves");
        const int64_t nn = n1 - n0, nl = l1 - l0;
        const int32_t base = (int32_t)hf->nodes.size();
        if (nn == 0) {
            hf->roots[t] = ~(int32_t)l0;  // the tree is the single leaf 0
            hf->tree_node_off.push_back((int64_t)hf->nodes.size());
            continue;
        }
        newidx.assign((size_t)nn, -1);
        order.clear();
        depth_of.assign((size_t)nn, 0);
        order.push_back(0);
        newidx[0] = base;
        depth_of[0] = 1;
        for (size_t qi = 0; qi < order.size(); ++qi) {
            const int32_t cur = order[qi];
            for (int b = 0; b < 2; ++b) {
                const int32_t c = raw.child[(size_t)(n0 + cur) * 2 + b];
                if (c >= 0) {
                    if (c >= nn) bad(DH_E_JSON, "child node index out of range in tree " + std::to_string(t));
                    if (newidx[c] != -1) bad(DH_E_JSON, "tree " + std::to_string(t) + " is not a tree (node reached twice: cycle or shared subtree)");
                    newidx[c] = base + (int32_t)order.size();
                    depth_of[c] = depth_of[cur] + 1;
                    order.push_back(c);
                } else {
                    const int64_t lf = (int64_t)(~c);
                    if (lf >= nl) bad(DH_E_JSON, "child leaf index out of range in tree " + std::to_string(t));
                }
            }
            max_depth = std::max(max_depth, depth_of[cur]);
        }
        for (int32_t old : order) {
            NodeRec rec;
            std::memset(&rec, 0, sizeof(rec));
            const int32_t* rc = &raw.rects[(size_t)(n0 + old) * 8];
            for (int k = 0; k < 2; ++k) {
                const int32_t x0 = rc[k * 4 + 0], y0 = rc[k * 4 + 1], x1 = rc[k * 4 + 2], y1 = rc[k * 4 + 3];
                if (x0 < 0 || y0 < 0 || x1 < x0 || y1 < y0)
                    bad(DH_E_JSON, "rectangle with bottomright < topleft (u32 underflow in Rect::width, types.rs:47-52)");
                if ((uint32_t)x1 > raw.subimage_width || (uint32_t)y1 > raw.subimage_height)
                    bad(DH_E_JSON, "feature rectangle exceeds the sub-image (the reference would read outside the patch)");
                rec.r[k * 4 + 0] = (uint8_t)x0;
                rec.r[k * 4 + 1] = (uint8_t)y0;
                rec.r[k * 4 + 2] = (uint8_t)x1;
                rec.r[k * 4 + 3] = (uint8_t)y1;
            }
            rec.threshold = raw.threshold[(size_t)(n0 + old)];
            {
                const uint32_t c1 = (uint32_t)(rec.r[2] - rec.r[0]) * (uint32_t)(rec.r[3] - rec.r[1]);
                const uint32_t c2 = (uint32_t)(rec.r[6] - rec.r[4]) * (uint32_t)(rec.r[7] - rec.r[5]);
                const double dd = (double)((uint64_t)(c1 ? c1 : 1u) * (uint64_t)(c2 ? c2 : 1u));  // exact (< 2^32)
                rec.thr_scaled = rec.threshold * dd;
            }
            for (int b = 0; b < 2; ++b) {
                const int32_t c = raw.child[(size_t)(n0 + old) * 2 + b];
                rec.child[b] = c >= 0 ? newidx[c] : ~(int32_t)(l0 + (int64_t)(~c));
            }
            hf->nodes.push_back(rec);
        }
        hf->roots[t] = base;
        hf->tree_node_off.push_back((int64_t)hf->nodes.size());
    }
    hf->max_depth = max_depth;
    if (!hf->nodes.empty()) {
        const NodeRec& n0r = hf->nodes[0];
        uint32_t rw = (uint32_t)(n0r.r[2] - n0r.r[0]), rh = (uint32_t)(n0r.r[3] - n0r.r[1]);
        for (const NodeRec& n : hf->nodes)
            for (int k = 0; k < 2; ++k)
                if ((uint32_t)(n.r[k * 4 + 2] - n.r[k * 4 + 0]) != rw || (uint32_t)(n.r[k * 4 + 3] - n.r[k * 4 + 1]) != rh) rw = rh = 0;
        // twice the difference of two rectangle sums must fit i32 (rw*rh*65535 < 2^30) for the integer node test
        if (rw && rh && rw * rh <= 16383u) {
            hf->uniform_rw = rw;
            hf->uniform_rh = rh;
        }
    }

    // ---- leaves and votes
    hf->leaf_prob = raw.prob;
    hf->leaf_vote_start.resize((size_t)NL);
    hf->leaf_n_votes.resize((size_t)NL);
    hf->offsets = raw.offsets;
    hf->rotations = raw.rotations;
    hf->rot_bins.resize((size_t)NV);
    uint32_t maxv = 0;
    for (int64_t l = 0; l < NL; ++l) {
        const int64_t v0 = raw.vote_off[l], v1 = raw.vote_off[l + 1];
        if (v1 < v0) bad(DH_E_ARG, "vote_off is not monotone");
        const int64_t n = v1 - v0;
        if (n > 0xFFFFF) bad(DH_E_JSON, "more than 1048575 votes in one leaf");
        const double p = raw.prob[l];
        if (std::isnan(p)) bad(DH_E_JSON, "leaf prob is NaN");
        if (p > 0.0 && n == 0)
            bad(DH_E_JSON, "leaf with prob > 0 and no votes: the reference divides by zero (prediction.rs:594)");
        hf->leaf_vote_start[l] = (uint32_t)v0;
        hf->leaf_n_votes[l] = (uint32_t)n;
        maxv = std::max(maxv, (uint32_t)n);
    }
    hf->max_votes_per_leaf = maxv;
    // rotation bins (prediction.rs:605-627): (rot*120.0/360.0) as i32 + 60, one conditional wrap.
    for (int64_t v = 0; v < NV; ++v) {
        uint32_t packed = 0;
        for (int k = 0; k < 3; ++k) {
            const double rv = raw.rotations[(size_t)v * 3 + k];
            const int32_t q = rs_f64_as_i32(rv * (double)kRotGridParts / 360.0);
            int64_t rr = (int64_t)q + kRotGridParts / 2;
            if (rr >= kRotGridParts) rr -= kRotGridParts;
            else if (rr < 0) rr += kRotGridParts;
            if (rr < 0 || rr >= kRotGridParts)
                bad(DH_E_JSON, "rotation vote outside (-540,540) degrees: the reference indexes outside its 20^3 "
                               "coarse grid (prediction.rs:630-636)");
            packed |= (uint32_t)rr << (8 * k);
        }
        hf->rot_bins[(size_t)v] = packed;
    }
    return hf.release();
}

void mat3_inverse_f32(const float m[9], float out[9]) {
    const float a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    const float adj[9] = {e * i - f * h, c * h - b * i, b * f - c * e, f * g - d * i, a * i - c * g,
                          c * d - a * f, d * h - e * g, b * g - a * h, a * e - b * d};
    // det (meancov_estimation.rs:339-343)
    const float det = a * (e * i - f * h) - d * (b * i - c * h) + g * (b * f - c * e);
    for (int k = 0; k < 9; ++k) out[k] = adj[k] / det;
}

void build_meanshift_kernel(float sigma, float* out) {
    const int n = kKernelSize, half = n / 2;
    for (int idx = 0; idx < n * n * n; ++idx) {
        const int z = idx / (n * n), rest = idx % (n * n), y = rest / n, x = rest % n;
        const int dx = x - half, dy = y - half, dz = z - half;
        const int norm = dx * dx + dy * dy + dz * dz;
        out[idx] = expf(-1.0f * (float)norm / (2.0f * sigma));  // kernel_function, meanshift.rs:228-232
    }
}

// imageproc 0.12 gaussian_kernel_f32 (the blur of build_hough_image, prediction.rs:844; external
// crate, restated from its published source: unpinned): radius ceil(2 sigma), taps
// (sqrt(2 pi) * sigma).recip() * exp(-i^2 / (2 sigma^2)) in f32, not normalised.
std::vector<float> build_gaussian_blur_kernel(float sigma) {
    if (!(sigma > 0.0f)) throw ModelError(DH_E_ARG, "gaussian blur: sigma must be > 0.0 (imageproc asserts it)");
    const float r2 = std::ceil(2.0f * sigma);
    if (!(r2 <= 1024.0f)) throw ModelError(DH_E_SHAPE, "gaussian blur: sigma above 512 is not supported");
    const size_t radius = (size_t)r2;
    std::vector<float> k(2 * radius + 1, 0.0f);
    const float two_pi = 2.0f * 3.14159265358979323846264338327950288f;
    const float norm = 1.0f / (std::sqrt(two_pi) * sigma);
    const float s2 = sigma * sigma;
    for (size_t i = 0; i <= radius; ++i) {
        const float x = (float)i, x2 = x * x;
        const float v = norm * std::exp(-x2 / (2.0f * s2));
        k[radius + i] = v;
        k[radius - i] = v;
    }
    return k;
}

}  // namespace dh

// ------------------------------------------------------------------------------------------------ writer
namespace dh {
namespace {
template <typename T>
void put_float(std::string& o, T v) {
    if (!std::isfinite(v)) { o += "null"; return; }   // serde_json writes non-finite floats as null
    char buf[40];
    auto r = std::to_chars(buf, buf + sizeof(buf), v);  // shortest representation that round-trips
    std::string t(buf, r.ptr);
    if (t.find_first_of(".eEn") == std::string::npos) t += ".0";  // serde_json/ryu always mark a float: 8 -> 8.0
    o += t;
}
void put_u(std::string& o, uint64_t v) { o += std::to_string(v); }
void put_rect(std::string& o, const uint8_t* r) {
    o += "{\"topleft\":["; put_u(o, r[0]); o += ','; put_u(o, r[1]);
    o += "],\"bottomright\":["; put_u(o, r[2]); o += ','; put_u(o, r[3]); o += "]}";
}
}  // namespace

std::string forest_to_json(const HostForest& hf) {
    std::string o;
    o.reserve(64 + hf.n_nodes() * 160 + hf.n_leaves() * 48 + hf.n_votes() * 90);
    o += "{\"stepwidth\":"; put_u(o, hf.stepwidth.load());
    o += ",\"subimage_width\":"; put_u(o, hf.subimage_width);
    o += ",\"subimage_height\":"; put_u(o, hf.subimage_height);
    o += ",\"gaussian_sigma\":"; put_float(o, hf.gaussian_sigma);
    o += ",\"forest\":{\"trees\":[";
    for (int32_t t = 0; t < hf.n_trees; ++t) {
        if (t) o += ',';
        const int64_t n0 = hf.tree_node_off[(size_t)t], n1 = hf.tree_node_off[(size_t)t + 1];
        const int64_t l0 = hf.tree_leaf_off[(size_t)t], l1 = hf.tree_leaf_off[(size_t)t + 1];
        o += "{\"functions\":{\"input_size\":{\"topleft\":[0,0],\"bottomright\":[";
        put_u(o, hf.subimage_width); o += ','; put_u(o, hf.subimage_height);
        o += "]},\"min_subrect_factor\":"; put_float(o, hf.fn_min_subrect_factor);
        o += ",\"max_subrect_factor\":"; put_float(o, hf.fn_max_subrect_factor);
        o += ",\"number_of_gen_features\":"; put_u(o, hf.fn_number_of_gen_features);
        o += ",\"steepness\":"; put_float(o, hf.fn_steepness);
        o += ",\"max_depth\":"; put_u(o, hf.fn_max_depth);
        o += ",\"min_subset_size\":"; put_u(o, hf.fn_min_subset_size);
        o += "},\"nodes\":[";
        for (int64_t n = n0; n < n1; ++n) {
            if (n > n0) o += ',';
            const NodeRec& r = hf.nodes[(size_t)n];
            o += "{\"param\":{\"r1\":"; put_rect(o, r.r);
            o += ",\"r2\":"; put_rect(o, r.r + 4);
            o += ",\"threshold\":"; put_float(o, r.threshold);
            o += "},\"children\":[";
            for (int b = 0; b < 2; ++b) {
                if (b) o += ',';
                const int32_t c = r.child[b];  // >= 0 global node, < 0 ~global leaf -> tree-local
                o += std::to_string(c >= 0 ? (int64_t)c - n0 : ~((int64_t)(~c) - l0));
            }
            o += "]}";
        }
        o += "],\"leaves\":[";
        for (int64_t l = l0; l < l1; ++l) {
            if (l > l0) o += ',';
            o += "{\"prob\":"; put_float(o, hf.leaf_prob[(size_t)l]);
            const uint32_t v0 = hf.leaf_vote_start[(size_t)l], nv = hf.leaf_n_votes[(size_t)l];
            o += ",\"offsets\":[";
            for (uint32_t v = 0; v < nv; ++v) {
                if (v) o += ',';
                o += '['; put_float(o, hf.offsets[(size_t)(v0 + v) * 3]); o += ','; put_float(o, hf.offsets[(size_t)(v0 + v) * 3 + 1]);
                o += ','; put_float(o, hf.offsets[(size_t)(v0 + v) * 3 + 2]); o += ']';
            }
            o += "],\"rotations\":[";
            for (uint32_t v = 0; v < nv; ++v) {
                if (v) o += ',';
                o += '['; put_float(o, hf.rotations[(size_t)(v0 + v) * 3]); o += ','; put_float(o, hf.rotations[(size_t)(v0 + v) * 3 + 1]);
                o += ','; put_float(o, hf.rotations[(size_t)(v0 + v) * 3 + 2]); o += ']';
            }
            o += "]}";
        }
        o += "]}";
    }
    o += "]},\"meanshift_iterations\":"; put_u(o, hf.meanshift_iterations.load());
    o += '}';
    return o;
}
}  // namespace dh
