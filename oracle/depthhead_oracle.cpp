// depthhead_oracle.cpp — CPU restatement of depthhead's Hough-forest prediction path.
//
// TEST INFRASTRUCTURE ONLY. Nothing under depthhead_b200/ (the product) may include, link,
// import or execute this file. Allowed callers: tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py — as the checker / reported CPU baseline,
// never as the thing shipped.
//
// Every function cites the reference lines it restates (paths relative to /root/reference).
// Build: see oracle/Makefile (-O2 -ffp-contract=off -fno-fast-math; Rust never fuses mul+add).
//
// PARITY STATUS
//   * pinned against the reference's own unit tests (types.rs:454-488,
//     meancov_estimation.rs:450-533) — see tests/test_oracle_kat.py;
//   * the forest container / root-to-leaf walk lives in the third-party crate stamm 0.2.0
//     (Cargo.lock:1154-1162), whose source is NOT under /root/reference, and the reference has
//     no test of prediction.rs / meanshift.rs / houghforest.rs.  For those parts this oracle is
//     "PARITY UNPINNED": it restates the visible call sites (prediction.rs:386,407;
//     houghforest.rs:185-193) with the traversal semantics documented in DESIGN.md
//     (child taken = children[bit], bit = avg1 - avg2 > threshold, leaves returned in tree order).
//
// Rust semantics replicated explicitly:
//   * `as i32` / `as usize` / `as u32` from floats: truncate toward zero, saturate, NaN -> 0;
//   * u32 vote sums wrap (release build);  * f32/f64 arithmetic is never contracted into FMA.

#include <algorithm>
#include <atomic>
#include <thread>
#include <sched.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------- Rust cast semantics
inline int32_t rs_f32_as_i32(float v) {
    if (std::isnan(v)) return 0;
    if (v >= 2147483648.0f) return INT32_MAX;
    if (v <= -2147483648.0f) return INT32_MIN;
    return (int32_t)v;  // C++ truncates toward zero for in-range values
}
inline int32_t rs_f64_as_i32(double v) {
    if (std::isnan(v)) return 0;
    if (v >= 2147483648.0) return INT32_MAX;
    if (v <= -2147483649.0) return INT32_MIN;
    return (int32_t)v;
}
inline uint64_t rs_f64_as_usize(double v) {
    if (std::isnan(v) || v <= 0.0) return 0;
    if (v >= 18446744073709551616.0) return UINT64_MAX;
    return (uint64_t)v;
}
inline uint64_t rs_f32_as_usize(float v) {
    if (std::isnan(v) || v <= 0.0f) return 0;
    if (v >= 18446744073709551616.0f) return UINT64_MAX;
    return (uint64_t)v;
}
inline uint8_t rs_f64_as_u8(double v) {
    if (std::isnan(v) || v <= 0.0) return 0;
    if (v >= 255.0) return 255;
    return (uint8_t)v;
}

// ---------------------------------------------------------------- types.rs:33-61  Rect
struct Rect {
    uint32_t tl[2];
    uint32_t br[2];
    uint32_t x() const { return tl[0]; }
    uint32_t y() const { return tl[1]; }
    uint32_t width() const { return br[0] - tl[0]; }   // u32 wrapping like a release build
    uint32_t height() const { return br[1] - tl[1]; }
};
inline Rect rect_new(uint32_t x, uint32_t y, uint32_t w, uint32_t h) {  // types.rs:40-45
    return Rect{{x, y}, {x + w, y + h}};
}
// types.rs:82-91
inline Rect rect_scale_and_replace(const Rect& r, double scale, double rx, double ry) {
    if (scale > 1.0) return r;
    double nw = (double)r.width() * scale;
    double nh = (double)r.height() * scale;
    double nx = (double)r.x() + rx * ((double)r.width() - nw);
    double ny = (double)r.y() + ry * ((double)r.height() - nh);
    auto as_u32 = [](double v) -> uint32_t {
        if (std::isnan(v) || v <= 0.0) return 0u;
        if (v >= 4294967295.0) return UINT32_MAX;
        return (uint32_t)v;
    };
    return rect_new(as_u32(nx), as_u32(ny), as_u32(nw), as_u32(nh));
}

// ---------------------------------------------------------------- meancov_estimation.rs
// Mat3 * Vec3 (meancov_estimation.rs:201-216): tmp = v0*m[j][0]; tmp = tmp + v_i*m[j][i]
template <typename T>
inline void mat3_mul_vec3(const T m[9], const T v[3], T out[3]) {
    for (int j = 0; j < 3; ++j) {
        T tmp = v[0] * m[j * 3 + 0];
        for (int i = 1; i < 3; ++i) tmp = tmp + v[i] * m[j * 3 + i];
        out[j] = tmp;
    }
}
// Mat3::det (meancov_estimation.rs:339-343)
template <typename T>
inline T mat3_det(const T m[9]) {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[3] * (m[1] * m[8] - m[2] * m[7]) +
           m[6] * (m[1] * m[5] - m[2] * m[4]);
}
// Mat3::inv (meancov_estimation.rs:344-352): adjugate / det, element-wise division
template <typename T>
inline void mat3_inv(const T m[9], T out[9]) {
    T a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], i = m[8];
    T adj[9] = {e * i - f * h, c * h - b * i, b * f - c * e, f * g - d * i, a * i - c * g,
                c * d - a * f, d * h - e * g, b * g - a * h, a * e - b * d};
    T det = mat3_det(m);
    for (int k = 0; k < 9; ++k) out[k] = adj[k] / det;
}
// Mat2 det / inv (meancov_estimation.rs:320-333)
template <typename T>
inline T mat2_det(const T m[4]) { return m[0] * m[3] - m[1] * m[2]; }
template <typename T>
inline void mat2_inv(const T m[4], T out[4]) {
    T det = mat2_det(m);
    T adj[4] = {m[3], -m[1], -m[2], m[0]};
    for (int k = 0; k < 4; ++k) out[k] = adj[k] / det;
}
// trace (meancov_estimation.rs:260-265): (0..n).map(..).sum() == ((0 + c00) + c11) + c22
template <typename T, int N>
inline T mat_trace(const T* m) {
    T s = (T)0;
    for (int i = 0; i < N; ++i) s = s + m[i * N + i];
    return s;
}
// estimate_mean_cov (meancov_estimation.rs:359-378).  Div<f64> for the f32 types casts the
// divisor to f32 first (meancov_estimation.rs:290-304); for f64 it is a plain division.
template <typename T, int N>
inline bool estimate_mean_cov(const T* set, size_t n, T* mean, T* cov) {
    if (n == 0) return false;
    for (int k = 0; k < N; ++k) mean[k] = set[k];
    for (size_t i = 1; i < n; ++i)
        for (int k = 0; k < N; ++k) mean[k] = mean[k] + set[i * N + k];
    T dn = (T)((double)n);
    for (int k = 0; k < N; ++k) mean[k] = mean[k] / dn;
    T d[N];
    for (int k = 0; k < N; ++k) d[k] = set[k] - mean[k];
    for (int a = 0; a < N; ++a)
        for (int b = 0; b < N; ++b) cov[a * N + b] = d[a] * d[b];
    for (size_t i = 1; i < n; ++i) {
        for (int k = 0; k < N; ++k) d[k] = set[i * N + k] - mean[k];
        for (int a = 0; a < N; ++a)
            for (int b = 0; b < N; ++b) cov[a * N + b] = cov[a * N + b] + d[a] * d[b];
    }
    T dn1 = (T)((double)(n - 1));
    for (int k = 0; k < N * N; ++k) cov[k] = cov[k] / dn1;
    return true;
}

// ---------------------------------------------------------------- types.rs:405-446 IntrinsicMatrix
struct Intrinsic {
    float k[9];
    float inv[9];
    bool have_inv = false;
};
// types.rs:424-428
inline void space_to_img_coord(const Intrinsic& m, const float p[3], float out[2]) {
    float res[3];
    mat3_mul_vec3<float>(m.k, p, res);
    float c = res[2];
    out[0] = res[0] / c;
    out[1] = res[1] / c;
}
// types.rs:432-445 (inverse cached lazily in the reference; same value either way)
inline void img_to_space_coord(Intrinsic& m, const float xy[2], float z, float out[3]) {
    if (!m.have_inv) {
        mat3_inv<float>(m.k, m.inv);
        m.have_inv = true;
    }
    float v3[3] = {xy[0], xy[1], 1.0f};
    float res[3];
    mat3_mul_vec3<float>(m.inv, v3, res);
    float c = z / res[2];
    out[0] = res[0] * c;
    out[1] = res[1] * c;
    out[2] = res[2] * c;
}

// ---------------------------------------------------------------- forest (stamm: UNPINNED)
struct Forest {
    int32_t n_trees = 0;
    std::vector<int64_t> tree_node_off;  // n_trees+1
    std::vector<int64_t> tree_leaf_off;  // n_trees+1
    std::vector<Rect> r1, r2;            // houghforest.rs:63-68 NodeParam
    std::vector<double> threshold;
    std::vector<int32_t> child;          // 2 per node; >=0 local node index, <0 = ~local leaf index
    std::vector<double> prob;            // houghforest.rs:73-78 LeafParam
    std::vector<int64_t> vote_off;       // n_leaves+1
    std::vector<float> offsets;          // 3 per vote
    std::vector<double> rotations;       // 3 per vote
};

// prediction.rs:239-256 (fields read by the prediction path)
struct Params {
    uint32_t stepwidth;
    uint32_t subimage_width;
    uint32_t subimage_height;
    float gaussian_sigma;
    uint32_t meanshift_iterations;
};

struct Image {
    const uint16_t* px;
    uint32_t w, h;
    uint16_t at(uint32_t x, uint32_t y) const { return px[(size_t)y * w + x]; }
};

// types.rs:317-339  SubImage::average_value_in_rect — the NAIVE loop, on purpose.
inline double average_value_in_rect(const Image& img, const Rect& sub, const Rect& rect) {
    uint64_t sum = 0, count = 0;
    for (uint32_t y = rect.y() + sub.y(); y < sub.y() + rect.y() + rect.height(); ++y)
        for (uint32_t x = rect.x() + sub.x(); x < sub.x() + rect.x() + rect.width(); ++x) {
            count += 1;
            sum += img.at(x, y);
        }
    if (count == 0) return 0.0;
    return (double)sum / (double)count;
}

// Summed-area variant (NOT in the reference) used by the "best-effort CPU" baseline and
// cross-checked against the naive loop in tests: sat is (h+1)x(w+1) u64, sat[y][x] = sum of
// pixels with x'<x, y'<y.
struct Sat {
    std::vector<uint64_t> s;
    uint32_t w = 0, h = 0;
    void build(const Image& img) {
        w = img.w;
        h = img.h;
        s.assign((size_t)(w + 1) * (h + 1), 0);
        for (uint32_t y = 0; y < h; ++y) {
            uint64_t row = 0;
            for (uint32_t x = 0; x < w; ++x) {
                row += img.at(x, y);
                s[(size_t)(y + 1) * (w + 1) + x + 1] = s[(size_t)y * (w + 1) + x + 1] + row;
            }
        }
    }
    uint64_t sum(uint32_t x0, uint32_t y0, uint32_t x1, uint32_t y1) const {
        const size_t p = w + 1;
        return s[(size_t)y1 * p + x1] - s[(size_t)y0 * p + x1] - s[(size_t)y1 * p + x0] +
               s[(size_t)y0 * p + x0];
    }
};
inline double average_value_in_rect_sat(const Sat& sat, const Rect& sub, const Rect& rect) {
    uint32_t x0 = rect.x() + sub.x(), y0 = rect.y() + sub.y();
    uint32_t x1 = x0 + rect.width(), y1 = y0 + rect.height();
    if (x1 <= x0 || y1 <= y0) return 0.0;
    uint64_t count = (uint64_t)(x1 - x0) * (y1 - y0);
    return (double)sat.sum(x0, y0, x1, y1) / (double)count;
}

// houghforest.rs:185-193  binarize: One iff avg1 - avg2 > threshold
template <typename AvgFn>
inline int binarize(const Forest& f, int64_t node, AvgFn&& avg) {
    double avg1 = avg(f.r1[node]);
    double avg2 = avg(f.r2[node]);
    return (avg1 - avg2 > f.threshold[node]) ? 1 : 0;
}

// stamm forest_predictions (UNPINNED): one leaf per tree, in tree order. Returns the GLOBAL
// leaf index (tree_leaf_off[t] + local).
template <typename AvgFn>
inline int64_t tree_predict(const Forest& f, int t, AvgFn&& avg, int* visited) {
    int64_t nbase = f.tree_node_off[t];
    int64_t nnodes = f.tree_node_off[t + 1] - nbase;
    int64_t lbase = f.tree_leaf_off[t];
    if (nnodes == 0) return lbase;  // tree is a single leaf
    int32_t cur = 0;
    int v = 0;
    for (;;) {
        ++v;
        int bit = binarize(f, nbase + cur, avg);
        int32_t c = f.child[(size_t)(nbase + cur) * 2 + bit];
        if (c < 0) {
            if (visited) *visited = v;
            return lbase + (int64_t)(~c);
        }
        cur = c;
    }
}

// ---------------------------------------------------------------- meanshift.rs
struct Key3 {
    int32_t x, y, z;
    bool operator==(const Key3& o) const { return x == o.x && y == o.y && z == o.z; }
    bool operator<(const Key3& o) const { return std::tie(x, y, z) < std::tie(o.x, o.y, o.z); }
};
struct Key3Hash {
    size_t operator()(const Key3& k) const {
        uint64_t h = 1469598103934665603ull;
        for (uint32_t v : {(uint32_t)k.x, (uint32_t)k.y, (uint32_t)k.z}) {
            h ^= v;
            h *= 1099511628211ull;
            h ^= h >> 29;
        }
        return (size_t)h;
    }
};
// meanshift.rs:14-68 SparseArray3D<u32>: default 0, insert-on-write
struct Sparse3D {
    std::unordered_map<Key3, uint32_t, Key3Hash> data;
    uint32_t get(int32_t x, int32_t y, int32_t z) const {
        auto it = data.find(Key3{x, y, z});
        return it == data.end() ? 0u : it->second;
    }
    void add(int32_t x, int32_t y, int32_t z, uint32_t v) { data[Key3{x, y, z}] += v; }  // wraps
};
// meanshift.rs:228-232 kernel_function ; :244-252 build_kernel ; dense index z*(w*h)+y*w+x (:78-88)
inline void build_kernel(uint32_t n, float variance, std::vector<float>& out) {
    out.resize((size_t)n * n * n);
    int32_t half = (int32_t)(n / 2);
    for (uint32_t i = 0; i < n * n * n; ++i) {
        uint32_t z = i / (n * n), rest = i % (n * n), y = rest / n, x = rest % n;
        int32_t dx = (int32_t)x - half, dy = (int32_t)y - half, dz = (int32_t)z - half;
        int32_t norm = dx * dx + dy * dy + dz * dz;
        out[i] = std::exp(-1.0f * (float)norm / (2.0f * variance));
    }
}
// meanshift.rs:328-407 for Idx = i32 (the `< i32::MIN` guards are no-ops).
// trace (optional): positions after every executed iteration, 3 ints each.
inline Key3 meanshift(const Sparse3D& acc, Key3 init, const std::vector<float>& kernel, int32_t ksz,
                      uint32_t iterations, std::vector<int32_t>* trace, int* zero_break) {
    int32_t pos[3] = {init.x, init.y, init.z};
    if (zero_break) *zero_break = 0;
    for (uint32_t it = 0; it < iterations; ++it) {
        float num[3] = {0.0f, 0.0f, 0.0f};
        float den = 0.0f;
        int32_t w = ksz, h = ksz, d = ksz;
        int32_t halfx = w / 2, halfy = h / 2, halfz = d / 2;
        for (int32_t x = -halfx; x < w - halfx; ++x)
            for (int32_t y = -halfy; y < h - halfy; ++y)
                for (int32_t z = -halfz; z < d - halfz; ++z) {
                    // wrapping add like a release build (positions never get near the limits)
                    int32_t ax = (int32_t)((uint32_t)pos[0] + (uint32_t)x);
                    int32_t ay = (int32_t)((uint32_t)pos[1] + (uint32_t)y);
                    int32_t az = (int32_t)((uint32_t)pos[2] + (uint32_t)z);
                    uint32_t factor = acc.get(ax, ay, az);
                    if (factor == 0) continue;
                    float influence = kernel[(size_t)(z + halfz) * (w * h) + (size_t)(y + halfy) * w +
                                             (size_t)(x + halfx)];
                    float fa[3] = {(float)ax, (float)ay, (float)az};
                    float ff = (float)factor;
                    float wgt = influence * ff;
                    for (int k = 0; k < 3; ++k) num[k] = num[k] + fa[k] * wgt;
                    den = den + influence * ff;
                }
        if (den == 0.0f) {
            if (zero_break) *zero_break = 1;
            break;
        }
        for (int k = 0; k < 3; ++k) pos[k] = rs_f32_as_i32(num[k] / den);
        if (trace) {
            trace->push_back(pos[0]);
            trace->push_back(pos[1]);
            trace->push_back(pos[2]);
        }
    }
    return Key3{pos[0], pos[1], pos[2]};
}

// ---------------------------------------------------------------- prediction.rs constants :270-286
constexpr uint32_t ZSCALEFACTOR = 1;
constexpr size_t GUESS_GRID_PARTS = 20;
constexpr size_t ROT_GRID_PARTS = 120;
constexpr double MAX_VARIANCE_ROT = 400.0;
constexpr float MAX_VARIANCE_OFFSET = 5200.0f;

// ---------------------------------------------------------------- fork-join pool
// Stand-in for the rayon pool behind stamm's forest_predictions_parallel (prediction.rs:407): one
// fork-join over the trees PER PATCH.  Persistent workers spin on an epoch counter (rayon's
// workers also spin before sleeping), grab tree indices from an atomic counter, and the caller
// joins by waiting for the done counter.  Two job slots alternate so that a straggler still
// looking at the previous job only ever sees an exhausted counter.
class TreePool {
public:
    typedef void (*Fn)(void* ctx, int index);
    explicit TreePool(int workers) {
        stop_.store(false);
        epoch_.store(0);
        for (int i = 0; i < 2; ++i) {
            slots_[i].next.store(1 << 30);
            slots_[i].done.store(0);
        }
        for (int i = 0; i < workers; ++i) threads_.emplace_back([this] { worker(); });
    }
    ~TreePool() {
        stop_.store(true);
        for (auto& t : threads_) t.join();
    }
    int workers() const { return (int)threads_.size(); }
    void run(int n, Fn fn, void* ctx) {
        const uint64_t e = epoch_.load(std::memory_order_relaxed) + 1;
        Job& j = slots_[e & 1];
        j.fn = fn;
        j.ctx = ctx;
        j.n = n;
        j.done.store(0, std::memory_order_relaxed);
        j.next.store(0, std::memory_order_release);
        epoch_.store(e, std::memory_order_release);
        drain(j);
        while (j.done.load(std::memory_order_acquire) < n) cpu_relax();
    }

private:
    struct Job {
        Fn fn = nullptr;
        void* ctx = nullptr;
        int n = 0;
        alignas(64) std::atomic<int> next;
        alignas(64) std::atomic<int> done;
    };
    static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
    static void drain(Job& j) {
        for (;;) {
            const int i = j.next.fetch_add(1, std::memory_order_acq_rel);
            if (i >= j.n) break;
            j.fn(j.ctx, i);
            j.done.fetch_add(1, std::memory_order_release);
        }
    }
    void worker() {
        uint64_t seen = 0;
        unsigned spins = 0;
        while (!stop_.load(std::memory_order_relaxed)) {
            const uint64_t e = epoch_.load(std::memory_order_acquire);
            if (e == seen) {
                cpu_relax();
                if (++spins > 20000) {
                    sched_yield();
                    spins = 0;
                }
                continue;
            }
            seen = e;
            spins = 0;
            drain(slots_[e & 1]);
        }
    }
    std::vector<std::thread> threads_;
    std::atomic<bool> stop_;
    alignas(64) std::atomic<uint64_t> epoch_;
    Job slots_[2];
};

enum Mode { MODE_NAIVE = 0, MODE_SAT = 1 };

struct Trace {
    // sliding window
    uint32_t npx = 0, npy = 0;
    std::vector<uint8_t> valid;       // P
    std::vector<int32_t> leaf;        // P*T, global leaf id, -1 = background patch
    std::vector<int32_t> visited;     // P*T, nodes visited (0 for background)
    std::vector<float> p3;            // P*3
    std::vector<uint8_t> gate;        // P: prob > 0.7
    std::vector<double> patch_prob;   // P
    // accumulators
    Sparse3D mid, rot;
    uint32_t guess_pos[GUESS_GRID_PARTS * GUESS_GRID_PARTS];
    std::vector<uint32_t> guess_rot;  // 8000, index z*400+y*20+x
    uint64_t n_mid_votes = 0, n_rot_votes = 0;
    // seeds / result
    int32_t seed_mid[3] = {0, 0, 0};
    double seed_rot_deg[3] = {0, 0, 0};
    int32_t seed_rot[3] = {0, 0, 0};
    std::vector<int32_t> ms_mid_trace, ms_rot_trace;
    int ms_mid_zero = 0, ms_rot_zero = 0;
    float mid_point[3] = {0, 0, 0};
    double rotation[3] = {0, 0, 0};
    int error = 0;
    std::string error_msg;
};

// prediction.rs:509-753 build_hough_cube_generic
void build_hough_cube(const Forest& forest, const Params& prm, const Image& img, Intrinsic& intr,
                      int mode, int tree_threads, bool keep_patch_trace, Trace& tr) {
    const uint32_t h = img.h, w = img.w;
    std::fill(std::begin(tr.guess_pos), std::end(tr.guess_pos), 0u);
    tr.guess_rot.assign(GUESS_GRID_PARTS * GUESS_GRID_PARTS * GUESS_GRID_PARTS, 0u);

    const uint32_t left_w = prm.subimage_width / 2;   // :535-538
    const uint32_t right_w = prm.subimage_width - left_w;
    const uint32_t left_h = prm.subimage_height / 2;
    const uint32_t right_h = prm.subimage_height - left_h;
    if (w < prm.subimage_width || h < prm.subimage_height || prm.stepwidth == 0) {
        tr.error = -2;
        tr.error_msg = "image smaller than the sub-image or zero step (reference underflows / never ends)";
        return;
    }
    Sat sat;
    if (mode == MODE_SAT) sat.build(img);

    const int T = forest.n_trees;
    tr.npx = 0;
    tr.npy = 0;
    for (uint32_t y = left_h; y < h - right_h; y += prm.stepwidth) tr.npy++;
    for (uint32_t x = left_w; x < w - right_w; x += prm.stepwidth) tr.npx++;
    const size_t P = (size_t)tr.npx * tr.npy;
    if (keep_patch_trace) {
        tr.valid.assign(P, 0);
        tr.leaf.assign(P * T, -1);
        tr.visited.assign(P * T, 0);
        tr.p3.assign(P * 3, 0.0f);
        tr.gate.assign(P, 0);
        tr.patch_prob.assign(P, 0.0);
    }
    std::vector<int64_t> leafs(T);
    std::vector<int> visited(T);
    std::unique_ptr<TreePool> pool;  // caller + (tree_threads-1) workers
    if (tree_threads > 1) pool.reset(new TreePool(tree_threads - 1));

    size_t pidx = 0;
    uint32_t y = left_h;  // :544-548
    while (y < h - right_h) {
        uint32_t x = left_w;
        while (x < w - right_w) {
            const uint16_t z = img.at(x, y);                              // :551
            float xy[2] = {(float)x, (float)y};
            float p3[3];
            img_to_space_coord(intr, xy, (float)z, p3);                  // :554
            const Rect sub = rect_new(x - left_w, y - left_h, prm.subimage_width, prm.subimage_height);
            const Rect whole = rect_new(0, 0, prm.subimage_width, prm.subimage_height);
            // :567-571 background test
            double bg = (mode == MODE_SAT) ? average_value_in_rect_sat(sat, sub, whole)
                                           : average_value_in_rect(img, sub, whole);
            bool have = bg > 0.0;
            if (have) {
                // :572 pred_func — forest_predictions{,_parallel}: trees are independent, the
                // returned Vec is in tree order either way.
                auto run_tree = [&](int t) {
                    int v = 0;
                    if (mode == MODE_SAT)
                        leafs[t] = tree_predict(forest, t, [&](const Rect& r) { return average_value_in_rect_sat(sat, sub, r); }, &v);
                    else
                        leafs[t] = tree_predict(forest, t, [&](const Rect& r) { return average_value_in_rect(img, sub, r); }, &v);
                    visited[t] = v;
                };
                if (pool) {
                    struct Ctx { decltype(run_tree)* f; } c{&run_tree};
                    pool->run(T, [](void* p, int t) { (*static_cast<Ctx*>(p)->f)(t); }, &c);
                } else {
                    for (int t = 0; t < T; ++t) run_tree(t);
                }
            }
            if (keep_patch_trace) {
                tr.valid[pidx] = have ? 1 : 0;
                tr.p3[pidx * 3 + 0] = p3[0];
                tr.p3[pidx * 3 + 1] = p3[1];
                tr.p3[pidx * 3 + 2] = p3[2];
                if (have)
                    for (int t = 0; t < T; ++t) {
                        tr.leaf[pidx * T + t] = (int32_t)leafs[t];
                        tr.visited[pidx * T + t] = visited[t];
                    }
            }
            if (have) {
                // :582 prob = sum(prob)/len — f64 fold from 0.0 in tree order
                double s = 0.0;
                for (int t = 0; t < T; ++t) s = s + forest.prob[leafs[t]];
                double prob = s / (double)T;
                if (keep_patch_trace) {
                    tr.patch_prob[pidx] = prob;
                    tr.gate[pidx] = prob > 0.7 ? 1 : 0;
                }
                if (prob > 0.7) {                                         // :584
                    for (int t = 0; t < T; ++t) {
                        const int64_t L = leafs[t];
                        const double lp = forest.prob[L];
                        if (!(lp > 0.0)) continue;                        // :590
                        const int64_t v0 = forest.vote_off[L], v1 = forest.vote_off[L + 1];
                        const size_t n = (size_t)(v1 - v0);
                        if (n == 0) {
                            tr.error = -3;
                            tr.error_msg = "leaf with prob>0 and no votes: reference divides by zero (prediction.rs:594)";
                            return;
                        }
                        uint64_t vta = rs_f64_as_usize(1000.0 * lp) / n;  // :594
                        uint32_t valtoadd = (uint32_t)vta;                // :595
                        // ---- rotation voting :600-638
                        double meanr[3], covr[9];
                        estimate_mean_cov<double, 3>(&forest.rotations[(size_t)v0 * 3], n, meanr, covr);
                        if (mat_trace<double, 3>(covr) <= MAX_VARIANCE_ROT) {
                            for (int64_t v = v0; v < v1; ++v) {
                                int32_t r[3];
                                for (int k = 0; k < 3; ++k) {
                                    double rv = forest.rotations[(size_t)v * 3 + k];
                                    // (rot * 120.0 / 360.0) as i32 + 60  — i32 add wraps in release
                                    int32_t q = rs_f64_as_i32(rv * (double)ROT_GRID_PARTS / 360.0);
                                    int32_t rr = (int32_t)((uint32_t)q + (uint32_t)(ROT_GRID_PARTS / 2));
                                    // inBetweenMod :616-627 — ONE conditional add/sub, not a modulo
                                    if (rr >= (int32_t)ROT_GRID_PARTS) rr = rr - (int32_t)ROT_GRID_PARTS;
                                    else if (rr < 0) rr = (int32_t)ROT_GRID_PARTS + rr;
                                    r[k] = rr;
                                }
                                uint32_t u[3] = {(uint32_t)r[0], (uint32_t)r[1], (uint32_t)r[2]};
                                uint32_t rough[3];
                                for (int k = 0; k < 3; ++k)
                                    rough[k] = u[k] * (uint32_t)GUESS_GRID_PARTS / (uint32_t)ROT_GRID_PARTS;  // u32 wrap mul
                                tr.rot.add((int32_t)u[0], (int32_t)u[1], (int32_t)u[2], valtoadd);  // :635
                                // :636 FullArray3D index (meanshift.rs:78-100): bounds check on the FLAT index only
                                size_t flat = (size_t)rough[2] * 400 + (size_t)rough[1] * 20 + (size_t)rough[0];
                                if (flat >= tr.guess_rot.size()) {
                                    tr.error = -4;
                                    tr.error_msg = "rotation vote outside the coarse grid: reference panics (prediction.rs:636)";
                                    return;
                                }
                                tr.guess_rot[flat] += valtoadd;
                                tr.n_rot_votes++;
                            }
                        }
                        // ---- head position voting :643-678
                        float meano[3], covo[9];
                        estimate_mean_cov<float, 3>(&forest.offsets[(size_t)v0 * 3], n, meano, covo);
                        if (mat_trace<float, 3>(covo) <= MAX_VARIANCE_OFFSET) {
                            for (int64_t v = v0; v < v1; ++v) {
                                float np[3];
                                for (int k = 0; k < 3; ++k) np[k] = p3[k] - forest.offsets[(size_t)v * 3 + k];  // :647
                                if (np[2] < 0.0f) continue;                                  // :650
                                float x3d = np[0], y3d = np[1], z3d = np[2];
                                float pt[3] = {x3d, y3d, z3d};
                                float p2[2];
                                space_to_img_coord(intr, pt, p2);                            // :661
                                // max!/min! macros :19-25 (plain comparisons; NaN -> 0.0)
                                float mx = (p2[0] > 0.0f) ? p2[0] : 0.0f;
                                float x2d = (mx < (float)(w - 1)) ? mx : (float)(w - 1);
                                float my = (p2[1] > 0.0f) ? p2[1] : 0.0f;
                                float y2d = (my < (float)(h - 1)) ? my : (float)(h - 1);
                                z3d = z3d / (float)ZSCALEFACTOR;                              // :666
                                tr.mid.add(rs_f32_as_i32(x3d), rs_f32_as_i32(y3d), rs_f32_as_i32(z3d), valtoadd);  // :667
                                size_t gx = (size_t)rs_f32_as_usize(x2d) * GUESS_GRID_PARTS / (size_t)w;  // :671
                                size_t gy = (size_t)rs_f32_as_usize(y2d) * GUESS_GRID_PARTS / (size_t)h;  // :673
                                tr.guess_pos[gy * GUESS_GRID_PARTS + gx] += valtoadd;          // :675
                                tr.n_mid_votes++;
                            }
                        }
                    }
                }
            }
            ++pidx;
            x += prm.stepwidth;  // :684
        }
        y += prm.stepwidth;      // :686
    }

    // :694-702 best 2d grid position: strict > from (0,0) -> first maximum, idx 0 if all zero
    uint32_t prev_max = 0;
    size_t best_idx = 0;
    for (size_t i = 0; i < GUESS_GRID_PARTS * GUESS_GRID_PARTS; ++i)
        if (tr.guess_pos[i] > prev_max) {
            prev_max = tr.guess_pos[i];
            best_idx = i;
        }
    // :706-729 z = mean of nonzero depth in the winning grid part
    size_t gpw = (size_t)w / GUESS_GRID_PARTS, gph = (size_t)h / GUESS_GRID_PARTS;
    size_t max_x_grid = best_idx % GUESS_GRID_PARTS, max_y_grid = best_idx / GUESS_GRID_PARTS;
    uint64_t zsum = 0;
    size_t zcnt = 0;
    for (size_t yy = 0; yy < gph; ++yy)
        for (size_t xx = 0; xx < gpw; ++xx) {
            uint16_t v = img.at((uint32_t)(gpw * max_x_grid + xx), (uint32_t)(gph * max_y_grid + yy));
            if (v > 0) {
                zsum += v;
                zcnt += 1;
            }
        }
    float meanz = zcnt > 0 ? (float)((double)zsum / (double)zcnt) : 0.0f;
    float mxy[2] = {((float)max_x_grid + 0.5f) * (float)gpw, ((float)max_y_grid + 0.5f) * (float)gph};
    float max3d[3];
    img_to_space_coord(intr, mxy, meanz, max3d);
    tr.seed_mid[0] = rs_f32_as_i32(max3d[0]);                                               // :750
    tr.seed_mid[1] = rs_f32_as_i32(max3d[1]);
    tr.seed_mid[2] = rs_f32_as_i32(max3d[2]) / (int32_t)ZSCALEFACTOR;

    // :733-747 best rotation guess: iteration order x fastest (meanshift.rs:114-138) == flat order
    uint32_t rx = 0, ry = 0, rz = 0, oldc = 0;
    for (size_t i = 0; i < tr.guess_rot.size(); ++i) {
        uint32_t c = tr.guess_rot[i];
        if (c > 0 && c > oldc) {
            oldc = c;
            rz = (uint32_t)(i / 400);
            ry = (uint32_t)((i % 400) / 20);
            rx = (uint32_t)(i % 20);
        }
    }
    tr.seed_rot_deg[0] = ((double)rx * 360.0 + 180.0) / (double)GUESS_GRID_PARTS;
    tr.seed_rot_deg[1] = ((double)ry * 360.0 + 180.0) / (double)GUESS_GRID_PARTS;
    tr.seed_rot_deg[2] = ((double)rz * 360.0 + 180.0) / (double)GUESS_GRID_PARTS;
}

// prediction.rs:421-493 predict_parameter_generic
void predict(const Forest& forest, const Params& prm, const Image& img, Intrinsic& intr,
             const float* midp_guess, const double* rot_guess, int mode, int tree_threads,
             bool keep_patch_trace, Trace& tr) {
    build_hough_cube(forest, prm, img, intr, mode, tree_threads, keep_patch_trace, tr);
    if (tr.error) return;
    if (midp_guess) {  // :437-441
        tr.seed_mid[0] = rs_f32_as_i32(midp_guess[0]);
        tr.seed_mid[1] = rs_f32_as_i32(midp_guess[1]);
        tr.seed_mid[2] = rs_f32_as_i32(midp_guess[2]) / (int32_t)ZSCALEFACTOR;
    }
    if (rot_guess)     // :444-453
        for (int k = 0; k < 3; ++k) tr.seed_rot_deg[k] = rot_guess[k] * 180.0 / 3.14159 + 180.0;
    for (int k = 0; k < 3; ++k)  // :458-460
        tr.seed_rot[k] = rs_f64_as_i32(tr.seed_rot_deg[k] * (double)ROT_GRID_PARTS / 360.0);

    std::vector<float> kernel;   // :310-317 get_or_build_kernel: build_kernel(20, gaussian_sigma)
    build_kernel(20, prm.gaussian_sigma, kernel);

    Key3 rm = meanshift(tr.mid, Key3{tr.seed_mid[0], tr.seed_mid[1], tr.seed_mid[2]}, kernel, 20,
                        prm.meanshift_iterations, &tr.ms_mid_trace, &tr.ms_mid_zero);   // :469
    Key3 rr = meanshift(tr.rot, Key3{tr.seed_rot[0], tr.seed_rot[1], tr.seed_rot[2]}, kernel, 20,
                        prm.meanshift_iterations, &tr.ms_rot_trace, &tr.ms_rot_zero);   // :472
    // :477-482
    const double half = (double)ROT_GRID_PARTS / 2.0, halfi = (double)(ROT_GRID_PARTS / 2);
    tr.rotation[0] = ((double)rr.x - half) / halfi * 3.14159;
    tr.rotation[1] = ((double)rr.y - half) / halfi * 3.14159;
    tr.rotation[2] = ((double)rr.z - half) / halfi * 3.14159;
    tr.mid_point[0] = (float)rm.x;  // :486-488
    tr.mid_point[1] = (float)rm.y;
    tr.mid_point[2] = (float)(int32_t)((uint32_t)rm.z * ZSCALEFACTOR);
}

// prediction.rs:850-905 predict_mask (uses forest_predictions, single core)
int predict_mask(const Forest& forest, const Params& prm, const Image& img, int mode, uint8_t* mask) {
    const uint32_t h = img.h, w = img.w;
    std::memset(mask, 0, (size_t)w * h);
    if (w < prm.subimage_width || h < prm.subimage_height || prm.stepwidth == 0) return -2;
    const uint32_t left_w = prm.subimage_width / 2, right_w = prm.subimage_width - left_w;
    const uint32_t left_h = prm.subimage_height / 2, right_h = prm.subimage_height - left_h;
    Sat sat;
    if (mode == MODE_SAT) sat.build(img);
    const int T = forest.n_trees;
    const uint32_t sw = prm.stepwidth;
    for (uint32_t y = left_h; y < h - right_h; y += sw)
        for (uint32_t x = left_w; x < w - right_w; x += sw) {
            const Rect sub = rect_new(x - left_w, y - left_h, prm.subimage_width, prm.subimage_height);
            const Rect whole = rect_new(0, 0, prm.subimage_width, prm.subimage_height);
            double bg = (mode == MODE_SAT) ? average_value_in_rect_sat(sat, sub, whole)
                                           : average_value_in_rect(img, sub, whole);
            if (!(bg > 0.0)) continue;
            double s = 0.0;
            for (int t = 0; t < T; ++t) {
                int64_t L = (mode == MODE_SAT)
                    ? tree_predict(forest, t, [&](const Rect& r) { return average_value_in_rect_sat(sat, sub, r); }, nullptr)
                    : tree_predict(forest, t, [&](const Rect& r) { return average_value_in_rect(img, sub, r); }, nullptr);
                s = s + forest.prob[L];
            }
            double prob = s / (double)T;                      // :881-882
            uint8_t pv = rs_f64_as_u8(prob * 255.0);          // :883
            for (uint32_t i = 0; i < sw; ++i)                 // :884-898
                for (uint32_t j = 0; j < sw; ++j) {
                    if (x + i < sw / 2 || y + j < sw / 2) continue;
                    if (x + i - sw / 2 >= w || y + j - sw / 2 >= h) continue;
                    mask[(size_t)(y + j - sw / 2) * w + (x + i - sw / 2)] = pv;
                }
        }
    return 0;
}

// prediction.rs:760-841 build_hough_image BEFORE the gaussian blur (:844 — imageproc is an
// external crate, blur parity unpinned; this returns the raw u16 vote image).
int hough_image_raw(const Forest& forest, const Params& prm, const Image& img, Intrinsic& intr, int mode,
                    uint16_t* out) {
    const uint32_t h = img.h, w = img.w;
    std::memset(out, 0, (size_t)w * h * sizeof(uint16_t));
    if (w < prm.subimage_width || h < prm.subimage_height || prm.stepwidth == 0) return -2;
    const uint32_t left_w = prm.subimage_width / 2, right_w = prm.subimage_width - left_w;
    const uint32_t left_h = prm.subimage_height / 2, right_h = prm.subimage_height - left_h;
    Sat sat;
    if (mode == MODE_SAT) sat.build(img);
    const int T = forest.n_trees;
    for (uint32_t y = left_h; y < h - right_h; y += prm.stepwidth)
        for (uint32_t x = left_w; x < w - right_w; x += prm.stepwidth) {
            uint16_t z = img.at(x, y);
            float xy[2] = {(float)x, (float)y}, p3[3];
            img_to_space_coord(intr, xy, (float)z, p3);
            const Rect sub = rect_new(x - left_w, y - left_h, prm.subimage_width, prm.subimage_height);
            const Rect whole = rect_new(0, 0, prm.subimage_width, prm.subimage_height);
            double bg = (mode == MODE_SAT) ? average_value_in_rect_sat(sat, sub, whole)
                                           : average_value_in_rect(img, sub, whole);
            if (!(bg > 0.0)) continue;
            for (int t = 0; t < T; ++t) {
                int64_t L = (mode == MODE_SAT)
                    ? tree_predict(forest, t, [&](const Rect& r) { return average_value_in_rect_sat(sat, sub, r); }, nullptr)
                    : tree_predict(forest, t, [&](const Rect& r) { return average_value_in_rect(img, sub, r); }, nullptr);
                double lp = forest.prob[L];
                if (!(lp >= 0.95)) continue;                                        // :805
                int64_t v0 = forest.vote_off[L], v1 = forest.vote_off[L + 1];
                if (v1 == v0) return -3;
                uint16_t valtoadd = (uint16_t)(rs_f64_as_usize(255.0 * lp) / (size_t)(v1 - v0));  // :807-808
                for (int64_t v = v0; v < v1; ++v) {
                    float np[3], p2[2];
                    for (int k = 0; k < 3; ++k) np[k] = p3[k] - forest.offsets[(size_t)v * 3 + k];
                    space_to_img_coord(intr, np, p2);
                    int32_t nx = rs_f32_as_i32(p2[0]), ny = rs_f32_as_i32(p2[1]);   // :816
                    if (nx < 0 || (uint32_t)nx >= w || ny < 0 || (uint32_t)ny >= h) continue;
                    out[(size_t)ny * w + nx] = (uint16_t)(out[(size_t)ny * w + nx] + valtoadd);  // u16 wrap
                }
            }
        }
    return 0;
}

}  // namespace

// =========================================================================== C API (ctypes)
extern "C" {

struct orc_forest { Forest f; };
struct orc_trace { Trace t; std::vector<int32_t> mid_keys, rot_keys; std::vector<uint32_t> mid_vals, rot_vals; };

// Forest arrays as produced by tests/_forest_py.py from the JSON document (independent of the
// product's C++ loader).  rects: int64[n_nodes][8] = r1.topleft x,y, r1.bottomright x,y, r2 ...
orc_forest* orc_forest_new(int32_t n_trees, const int64_t* tree_node_off, const int64_t* tree_leaf_off,
                           const int64_t* rects, const double* threshold, const int32_t* child,
                           const double* prob, const int64_t* vote_off, const float* offsets,
                           const double* rotations) {
    auto* o = new orc_forest();
    Forest& f = o->f;
    f.n_trees = n_trees;
    f.tree_node_off.assign(tree_node_off, tree_node_off + n_trees + 1);
    f.tree_leaf_off.assign(tree_leaf_off, tree_leaf_off + n_trees + 1);
    int64_t nn = f.tree_node_off[n_trees], nl = f.tree_leaf_off[n_trees];
    f.r1.resize(nn);
    f.r2.resize(nn);
    for (int64_t i = 0; i < nn; ++i) {
        const int64_t* r = rects + i * 8;
        f.r1[i] = Rect{{(uint32_t)r[0], (uint32_t)r[1]}, {(uint32_t)r[2], (uint32_t)r[3]}};
        f.r2[i] = Rect{{(uint32_t)r[4], (uint32_t)r[5]}, {(uint32_t)r[6], (uint32_t)r[7]}};
    }
    f.threshold.assign(threshold, threshold + nn);
    f.child.assign(child, child + nn * 2);
    f.prob.assign(prob, prob + nl);
    f.vote_off.assign(vote_off, vote_off + nl + 1);
    int64_t nv = f.vote_off[nl];
    f.offsets.assign(offsets, offsets + nv * 3);
    f.rotations.assign(rotations, rotations + nv * 3);
    return o;
}
void orc_forest_free(orc_forest* f) { delete f; }

orc_trace* orc_trace_new() { return new orc_trace(); }
void orc_trace_free(orc_trace* t) { delete t; }

static void dump_sorted(const Sparse3D& s, std::vector<int32_t>& keys, std::vector<uint32_t>& vals) {
    std::vector<std::pair<Key3, uint32_t>> v(s.data.begin(), s.data.end());
    std::sort(v.begin(), v.end(), [](auto& a, auto& b) { return a.first < b.first; });
    keys.clear();
    vals.clear();
    for (auto& kv : v) {
        keys.push_back(kv.first.x);
        keys.push_back(kv.first.y);
        keys.push_back(kv.first.z);
        vals.push_back(kv.second);
    }
}

// mode: 0 naive rect sums (faithful), 1 summed-area table.  tree_threads: >1 evaluates the trees
// of one patch in parallel (the shape of forest_predictions_parallel).  keep: patch-level trace.
int orc_predict(const orc_forest* forest, uint32_t stepwidth, uint32_t sub_w, uint32_t sub_h, float sigma,
                uint32_t iterations, const uint16_t* depth, uint32_t w, uint32_t h, const float* K,
                const float* midp_guess, const double* rot_guess, int mode, int tree_threads, int keep,
                orc_trace* out) {
    out->t = Trace();
    Params prm{stepwidth, sub_w, sub_h, sigma, iterations};
    Image img{depth, w, h};
    Intrinsic intr;
    std::memcpy(intr.k, K, sizeof(float) * 9);
    predict(forest->f, prm, img, intr, midp_guess, rot_guess, mode, tree_threads, keep != 0, out->t);
    if (keep) {
        dump_sorted(out->t.mid, out->mid_keys, out->mid_vals);
        dump_sorted(out->t.rot, out->rot_keys, out->rot_vals);
    }
    return out->t.error;
}
const char* orc_trace_error(const orc_trace* t) { return t->t.error_msg.c_str(); }
void orc_trace_result(const orc_trace* t, float* mid_point, double* rotation) {
    std::memcpy(mid_point, t->t.mid_point, sizeof(float) * 3);
    std::memcpy(rotation, t->t.rotation, sizeof(double) * 3);
}
void orc_trace_seeds(const orc_trace* t, int32_t* seed_mid, int32_t* seed_rot, double* seed_rot_deg) {
    std::memcpy(seed_mid, t->t.seed_mid, sizeof(int32_t) * 3);
    std::memcpy(seed_rot, t->t.seed_rot, sizeof(int32_t) * 3);
    std::memcpy(seed_rot_deg, t->t.seed_rot_deg, sizeof(double) * 3);
}
void orc_trace_dims(const orc_trace* t, uint32_t* npx, uint32_t* npy, uint64_t* n_mid_votes, uint64_t* n_rot_votes,
                    uint64_t* n_mid_cells, uint64_t* n_rot_cells, uint32_t* n_ms_mid, uint32_t* n_ms_rot,
                    int32_t* ms_mid_zero, int32_t* ms_rot_zero) {
    *npx = t->t.npx;
    *npy = t->t.npy;
    *n_mid_votes = t->t.n_mid_votes;
    *n_rot_votes = t->t.n_rot_votes;
    *n_mid_cells = t->t.mid.data.size();
    *n_rot_cells = t->t.rot.data.size();
    *n_ms_mid = (uint32_t)(t->t.ms_mid_trace.size() / 3);
    *n_ms_rot = (uint32_t)(t->t.ms_rot_trace.size() / 3);
    *ms_mid_zero = t->t.ms_mid_zero;
    *ms_rot_zero = t->t.ms_rot_zero;
}
void orc_trace_patches(const orc_trace* t, uint8_t* valid, int32_t* leaf, int32_t* visited, float* p3,
                       uint8_t* gate, double* patch_prob) {
    const Trace& tr = t->t;
    if (valid) std::memcpy(valid, tr.valid.data(), tr.valid.size());
    if (leaf) std::memcpy(leaf, tr.leaf.data(), tr.leaf.size() * sizeof(int32_t));
    if (visited) std::memcpy(visited, tr.visited.data(), tr.visited.size() * sizeof(int32_t));
    if (p3) std::memcpy(p3, tr.p3.data(), tr.p3.size() * sizeof(float));
    if (gate) std::memcpy(gate, tr.gate.data(), tr.gate.size());
    if (patch_prob) std::memcpy(patch_prob, tr.patch_prob.data(), tr.patch_prob.size() * sizeof(double));
}
void orc_trace_grids(const orc_trace* t, uint32_t* guess_pos, uint32_t* guess_rot) {
    std::memcpy(guess_pos, t->t.guess_pos, sizeof(t->t.guess_pos));
    std::memcpy(guess_rot, t->t.guess_rot.data(), t->t.guess_rot.size() * sizeof(uint32_t));
}
// accumulators as key-sorted lists: keys int32[n][3], vals uint32[n]
void orc_trace_accumulators(const orc_trace* t, int32_t* mid_keys, uint32_t* mid_vals, int32_t* rot_keys,
                            uint32_t* rot_vals) {
    if (mid_keys) std::memcpy(mid_keys, t->mid_keys.data(), t->mid_keys.size() * sizeof(int32_t));
    if (mid_vals) std::memcpy(mid_vals, t->mid_vals.data(), t->mid_vals.size() * sizeof(uint32_t));
    if (rot_keys) std::memcpy(rot_keys, t->rot_keys.data(), t->rot_keys.size() * sizeof(int32_t));
    if (rot_vals) std::memcpy(rot_vals, t->rot_vals.data(), t->rot_vals.size() * sizeof(uint32_t));
}
void orc_trace_meanshift(const orc_trace* t, int32_t* ms_mid, int32_t* ms_rot) {
    if (ms_mid) std::memcpy(ms_mid, t->t.ms_mid_trace.data(), t->t.ms_mid_trace.size() * sizeof(int32_t));
    if (ms_rot) std::memcpy(ms_rot, t->t.ms_rot_trace.data(), t->t.ms_rot_trace.size() * sizeof(int32_t));
}

// Batch driver for the CPU baseline: n frames, results only. frame_threads>1 = the "best-effort
// CPU" line (frame-level OpenMP, use with mode=1); tree_threads>1 = the reference's shape.
int orc_predict_batch(const orc_forest* forest, uint32_t stepwidth, uint32_t sub_w, uint32_t sub_h, float sigma,
                      uint32_t iterations, const uint16_t* depth, uint32_t n, uint32_t w, uint32_t h,
                      const float* K, int mode, int tree_threads, int frame_threads, float* mid_points,
                      double* rotations, uint64_t* evals_out) {
    int err = 0;
    uint64_t evals = 0;
#ifdef _OPENMP
#pragma omp parallel for num_threads(frame_threads > 1 ? frame_threads : 1) schedule(dynamic) reduction(+ : evals) if (frame_threads > 1)
#endif
    for (int64_t i = 0; i < (int64_t)n; ++i) {
        Trace tr;
        Params prm{stepwidth, sub_w, sub_h, sigma, iterations};
        Image img{depth + (size_t)i * w * h, w, h};
        Intrinsic intr;
        std::memcpy(intr.k, K, sizeof(float) * 9);
        predict(forest->f, prm, img, intr, nullptr, nullptr, mode, frame_threads > 1 ? 1 : tree_threads, true, tr);
        if (tr.error) {
#ifdef _OPENMP
#pragma omp critical
#endif
            err = tr.error;
            continue;
        }
        uint64_t nv = 0;
        for (uint8_t v : tr.valid) nv += v;
        evals += nv * (uint64_t)forest->f.n_trees;
        std::memcpy(mid_points + i * 3, tr.mid_point, sizeof(float) * 3);
        std::memcpy(rotations + i * 3, tr.rotation, sizeof(double) * 3);
    }
    if (evals_out) *evals_out = evals;
    return err;
}

int orc_predict_mask(const orc_forest* forest, uint32_t stepwidth, uint32_t sub_w, uint32_t sub_h,
                     const uint16_t* depth, uint32_t w, uint32_t h, int mode, uint8_t* mask) {
    Params prm{stepwidth, sub_w, sub_h, 1.0f, 0};
    Image img{depth, w, h};
    return predict_mask(forest->f, prm, img, mode, mask);
}
int orc_hough_image_raw(const orc_forest* forest, uint32_t stepwidth, uint32_t sub_w, uint32_t sub_h,
                        const uint16_t* depth, uint32_t w, uint32_t h, const float* K, int mode, uint16_t* out) {
    Params prm{stepwidth, sub_w, sub_h, 1.0f, 0};
    Image img{depth, w, h};
    Intrinsic intr;
    std::memcpy(intr.k, K, sizeof(float) * 9);
    return hough_image_raw(forest->f, prm, img, intr, mode, out);
}

// ---- build_hough_image's blur and predict_parameter_from2dhough (prediction.rs:343-367, 841-845)
// PARITY UNPINNED: gaussian_blur_f32 lives in the external crate imageproc 0.12 (Cargo.toml:26), whose
// source is not under /root/reference.  Restated here from its published implementation
// (imageproc/src/filter.rs: gaussian_blur_f32 -> gaussian_kernel_f32 -> separable_filter_equal ->
// horizontal_filter, vertical_filter; definitions.rs: Clamp<f32> for u16):
//   * kernel: radius = ceil(2 * sigma) taps on either side, k[r +- i] = gaussian(i as f32, sigma) with
//     gaussian(x, r) = (sqrt(2 pi) * r).recip() * exp(-x^2 / (2 r^2)), all f32, NOT normalised;
//   * horizontal pass over the u16 image, then vertical pass over ITS u16 result; per output pixel
//     acc = 0.0f32; for the taps in order: acc = acc + (pixel as f32) * k[i] (never fused);
//   * border rule: the coordinate is clamped to the image (edge pixels repeat);
//   * every pass stores Clamp<f32>::clamp(acc) as u16: >= 65535 -> 65535, <= 0 -> 0, else truncation.
static inline float gaussian_pdf_f32(float x, float r) {
    const float two_pi = 2.0f * 3.14159265358979323846264338327950288f;  // 2.0 * f32::consts::PI
    const float norm = 1.0f / (std::sqrt(two_pi) * r);                      // .recip()
    const float r2 = r * r, x2 = x * x;                                     // powi(2)
    return norm * std::exp(-x2 / (2.0f * r2));
}
static std::vector<float> gaussian_kernel_f32(float sigma) {
    const size_t radius = (size_t)std::ceil(2.0f * sigma);
    std::vector<float> k(2 * radius + 1, 0.0f);
    for (size_t i = 0; i <= radius; ++i) {
        const float v = gaussian_pdf_f32((float)i, sigma);
        k[radius + i] = v;
        k[radius - i] = v;
    }
    return k;
}
static inline uint16_t clamp_f32_to_u16(float x) {
    if (x < 65535.0f) return x > 0.0f ? (uint16_t)x : (uint16_t)0;
    return 65535;  // also NaN, as the comparison chain of the crate does
}
static void gaussian_blur_u16(const uint16_t* in, uint32_t w, uint32_t h, float sigma, uint16_t* out) {
    const std::vector<float> k = gaussian_kernel_f32(sigma);
    const int kw = (int)k.size(), half = kw / 2;
    std::vector<uint16_t> tmp((size_t)w * h);
    for (uint32_t y = 0; y < h; ++y)
        for (uint32_t x = 0; x < w; ++x) {
            float acc = 0.0f;
            for (int i = 0; i < kw; ++i) {
                const int xu = (int)x + i - half;
                const uint32_t xp = (uint32_t)std::min<int>(std::max(xu, 0), (int)w - 1);
                acc = acc + (float)in[(size_t)y * w + xp] * k[(size_t)i];
            }
            tmp[(size_t)y * w + x] = clamp_f32_to_u16(acc);
        }
    for (uint32_t y = 0; y < h; ++y)
        for (uint32_t x = 0; x < w; ++x) {
            float acc = 0.0f;
            for (int i = 0; i < kw; ++i) {
                const int yu = (int)y + i - half;
                const uint32_t yp = (uint32_t)std::min<int>(std::max(yu, 0), (int)h - 1);
                acc = acc + (float)tmp[(size_t)yp * w + x] * k[(size_t)i];
            }
            out[(size_t)y * w + x] = clamp_f32_to_u16(acc);
        }
}
int orc_gaussian_kernel_f32(float sigma, float* out, uint32_t cap) {
    const std::vector<float> k = gaussian_kernel_f32(sigma);
    if (out && cap >= k.size()) std::memcpy(out, k.data(), k.size() * sizeof(float));
    return (int)k.size();
}
void orc_gaussian_blur_u16(const uint16_t* in, uint32_t w, uint32_t h, float sigma, uint16_t* out) {
    gaussian_blur_u16(in, w, h, sigma, out);
}
// build_hough_image (prediction.rs:760-845): the vote image, then the blur with gaussian_sigma
int orc_build_hough_image(const orc_forest* forest, uint32_t stepwidth, uint32_t sub_w, uint32_t sub_h, float sigma,
                          const uint16_t* depth, uint32_t w, uint32_t h, const float* K, int mode, uint16_t* out) {
    std::vector<uint16_t> raw((size_t)w * h);
    const int rc = orc_hough_image_raw(forest, stepwidth, sub_w, sub_h, depth, w, h, K, mode, raw.data());
    if (rc) return rc;
    if (!(sigma > 0.0f)) return 7;  // the crate asserts sigma > 0.0
    gaussian_blur_u16(raw.data(), w, h, sigma, out);
    return 0;
}
// predict_parameter_from2dhough (prediction.rs:343-367): arg-max of the blurred vote image with
// Iterator::max_by_key (the LAST of equal maxima wins), z = the depth at that pixel,
// mid_point = img_to_space_coord([x, y], z), rotation 0
int orc_predict_from2dhough(const orc_forest* forest, uint32_t stepwidth, uint32_t sub_w, uint32_t sub_h, float sigma,
                            const uint16_t* depth, uint32_t w, uint32_t h, const float* K, int mode, float* mid_point,
                            uint32_t* best_xy) {
    if ((uint64_t)w * h == 0) return 8;  // max_by_key on an empty range: unwrap of None
    std::vector<uint16_t> img((size_t)w * h);
    const int rc = orc_build_hough_image(forest, stepwidth, sub_w, sub_h, sigma, depth, w, h, K, mode, img.data());
    if (rc) return rc;
    size_t best = 0;
    for (size_t i = 1; i < img.size(); ++i)
        if (img[i] >= img[best]) best = i;
    const uint32_t x = (uint32_t)(best % w), y = (uint32_t)(best / w);
    Intrinsic m;
    std::memcpy(m.k, K, sizeof(float) * 9);
    const float xy[2] = {(float)x, (float)y};
    img_to_space_coord(m, xy, (float)depth[(size_t)y * w + x], mid_point);
    if (best_xy) {
        best_xy[0] = x;
        best_xy[1] = y;
    }
    return 0;
}

// ---- small pieces exposed for the known-answer tests (reference unit tests)
void orc_rect_scale_and_replace(const uint32_t* xywh, double scale, double rx, double ry, uint32_t* out_xywh) {
    Rect r = rect_new(xywh[0], xywh[1], xywh[2], xywh[3]);
    Rect o = rect_scale_and_replace(r, scale, rx, ry);
    out_xywh[0] = o.x();
    out_xywh[1] = o.y();
    out_xywh[2] = o.width();
    out_xywh[3] = o.height();
}
void orc_space_to_img(const float* K, const float* p3, float* out2) {
    Intrinsic m;
    std::memcpy(m.k, K, sizeof(float) * 9);
    space_to_img_coord(m, p3, out2);
}
void orc_img_to_space(const float* K, const float* xy, float z, float* out3) {
    Intrinsic m;
    std::memcpy(m.k, K, sizeof(float) * 9);
    img_to_space_coord(m, xy, z, out3);
}
void orc_mat3_inv_f32(const float* m, float* out) { mat3_inv<float>(m, out); }
void orc_mat3_inv_f64(const double* m, double* out) { mat3_inv<double>(m, out); }
double orc_mat3_det_f64(const double* m) { return mat3_det<double>(m); }
double orc_mat3_trace_f64(const double* m) { return mat_trace<double, 3>(m); }
double orc_mat2_det_f64(const double* m) { return mat2_det<double>(m); }
double orc_mat2_trace_f64(const double* m) { return mat_trace<double, 2>(m); }
void orc_mat2_inv_f64(const double* m, double* out) { mat2_inv<double>(m, out); }
void orc_mat3_mul_vec3_f64(const double* m, const double* v, double* out) { mat3_mul_vec3<double>(m, v, out); }
void orc_mat2_mul_vec2_f64(const double* m, const double* v, double* out) {
    for (int j = 0; j < 2; ++j) {
        double tmp = v[0] * m[j * 2 + 0];
        tmp = tmp + v[1] * m[j * 2 + 1];
        out[j] = tmp;
    }
}
int orc_mean_cov3_f64(const double* set, uint64_t n, double* mean, double* cov) {
    return estimate_mean_cov<double, 3>(set, n, mean, cov) ? 0 : -1;
}
int orc_mean_cov3_f32(const float* set, uint64_t n, float* mean, float* cov) {
    return estimate_mean_cov<float, 3>(set, n, mean, cov) ? 0 : -1;
}
int orc_mean_cov2_f64(const double* set, uint64_t n, double* mean, double* cov) {
    return estimate_mean_cov<double, 2>(set, n, mean, cov) ? 0 : -1;
}
void orc_build_kernel(uint32_t n, float variance, float* out) {
    std::vector<float> k;
    build_kernel(n, variance, k);
    std::memcpy(out, k.data(), k.size() * sizeof(float));
}
// average_value_in_rect, naive vs SAT, for one (sub, rect) pair
double orc_rect_average(const uint16_t* depth, uint32_t w, uint32_t h, const uint32_t* sub_xywh,
                        const uint32_t* rect_tlbr, int mode) {
    Image img{depth, w, h};
    Rect sub = rect_new(sub_xywh[0], sub_xywh[1], sub_xywh[2], sub_xywh[3]);
    Rect r{{rect_tlbr[0], rect_tlbr[1]}, {rect_tlbr[2], rect_tlbr[3]}};
    if (mode == MODE_SAT) {
        Sat sat;
        sat.build(img);
        return average_value_in_rect_sat(sat, sub, r);
    }
    return average_value_in_rect(img, sub, r);
}
// leaf-static quantities (prediction.rs:594-600,643) for one leaf — used to check the CUDA
// leaf-gate kernel: valtoadd, rot trace (f64), offset trace (f32)
void orc_leaf_static(const orc_forest* forest, int64_t leaf, uint32_t* valtoadd, double* trace_rot,
                     float* trace_off) {
    const Forest& f = forest->f;
    int64_t v0 = f.vote_off[leaf], v1 = f.vote_off[leaf + 1];
    size_t n = (size_t)(v1 - v0);
    if (n == 0) {
        *valtoadd = 0;
        *trace_rot = std::numeric_limits<double>::quiet_NaN();
        *trace_off = std::numeric_limits<float>::quiet_NaN();
        return;
    }
    *valtoadd = (uint32_t)(rs_f64_as_usize(1000.0 * f.prob[leaf]) / n);
    double mr[3], cr[9];
    estimate_mean_cov<double, 3>(&f.rotations[(size_t)v0 * 3], n, mr, cr);
    *trace_rot = mat_trace<double, 3>(cr);
    float mo[3], co[9];
    estimate_mean_cov<float, 3>(&f.offsets[(size_t)v0 * 3], n, mo, co);
    *trace_off = mat_trace<float, 3>(co);
}
int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// ---------------------------------------------------------------------------- Biwi wire formats
// src/db_reader/biwi.rs restated.  Return codes: 0 ok, 1 the reader ran out of bytes (io::Error
// UnexpectedEof in the reference), 2 a run writes past the last pixel (`it.next().unwrap()` panics,
// biwi.rs:92,97), 3 caller's buffer too small (not a reference condition).
struct ByteReader {  // byteorder::ReadBytesExt over a slice, little endian
    const uint8_t* p;
    size_t len, pos;
    bool u32(uint32_t* v) {
        if (pos + 4 > len) { pos = len; return false; }
        *v = (uint32_t)p[pos] | ((uint32_t)p[pos + 1] << 8) | ((uint32_t)p[pos + 2] << 16) | ((uint32_t)p[pos + 3] << 24);
        pos += 4;
        return true;
    }
    bool u16(uint16_t* v) {
        if (pos + 2 > len) { pos = len; return false; }
        *v = (uint16_t)(p[pos] | (p[pos + 1] << 8));
        pos += 2;
        return true;
    }
};

// read_depth, biwi.rs:81-103
int orc_biwi_read_depth(const uint8_t* file, uint64_t len, uint16_t* out, uint64_t cap_px, uint32_t* w_out, uint32_t* h_out) {
    ByteReader r{file, (size_t)len, 0};
    uint32_t width, height;
    if (!r.u32(&width)) return 1;                       // :83
    if (!r.u32(&height)) return 1;                      // :84
    *w_out = width;
    *h_out = height;
    const uint64_t npx = (uint64_t)(uint32_t)(width * height);  // `(width * height) as usize`: u32 product (:89)
    if (npx > cap_px) return 3;
    for (uint64_t i = 0; i < npx; ++i) out[i] = 0;      // DepthImage::new (:86)
    uint64_t p = 0, it = 0;                             // it: pixels_mut() iterator position
    while (p < npx) {                                   // :89
        uint32_t num_empty, num_nonempty;
        if (!r.u32(&num_empty)) return 1;               // :90
        for (uint32_t k = 0; k < num_empty; ++k) {      // :91-93
            if (it >= npx) return 2;
            out[it++] = 0;
        }
        if (!r.u32(&num_nonempty)) return 1;            // :94
        for (uint32_t k = 0; k < num_nonempty; ++k) {   // :95-98
            uint16_t v;
            if (!r.u16(&v)) return 1;
            if (it >= npx) return 2;
            out[it++] = v;
        }
        p += (uint64_t)num_empty + (uint64_t)num_nonempty;  // :99
    }
    return 0;
}

// read_gt, biwi.rs:63-77 (+ space_to_img_coord, types.rs:424-428)
int orc_biwi_read_gt(const uint8_t* file, uint64_t len, const float* K, float* pos3d, float* pos2d, float* rot) {
    ByteReader r{file, (size_t)len, 0};
    float res[6];
    for (int i = 0; i < 6; ++i) {
        uint32_t u;
        if (!r.u32(&u)) return 1;
        std::memcpy(&res[i], &u, 4);
    }
    Intrinsic k;
    std::memcpy(k.k, K, sizeof(float) * 9);
    const float p3[3] = {res[0], res[1], res[2]};
    float p2[2];
    space_to_img_coord(k, p3, p2);
    for (int i = 0; i < 3; ++i) {
        pos3d[i] = res[i];
        rot[i] = res[3 + i];
    }
    pos2d[0] = p2[0];
    pos2d[1] = p2[1];
    return 0;
}

// read_cal, biwi.rs:27-60.  Return: 0 ok, 1 "Unsupported Calibration-File", 2 ParseFloatError,
// 3 a fourth match on a line (index out of bounds panic at :45).
int orc_biwi_read_cal(const char* text, uint64_t len, float* K) {
    auto digit = [](char c) { return c >= '0' && c <= '9'; };
    size_t pos = 0;
    for (int j = 0; j < 3; ++j) {
        size_t e = pos;
        while (e < len && text[e] != '\n') ++e;          // read_line
        int found = 0;
        size_t i = pos;
        while (i < e) {                                  // FLOAT.captures_iter: (\d+[\.\d+]*)
            if (!digit(text[i])) { ++i; continue; }
            size_t k = i;
            while (k < e && digit(text[k])) ++k;
            while (k < e && (digit(text[k]) || text[k] == '.' || text[k] == '+')) ++k;
            if (found == 3) return 3;                    // res[j][3] = ..  (:45)
            // f32::from_str: digits [ '.' digits ] for a token made of these characters
            std::string tok(text + i, k - i);
            size_t d = 0;
            while (d < tok.size() && digit(tok[d])) ++d;
            if (d < tok.size() && tok[d] == '.') { ++d; while (d < tok.size() && digit(tok[d])) ++d; }
            if (d != tok.size()) return 2;
            K[j * 3 + found] = strtof(tok.c_str(), nullptr);
            ++found;
            i = k;
        }
        if (found != 3) return 1;                        // :53-55
        pos = e < len ? e + 1 : len;
    }
    return 0;
}

// ---------------------------------------------------------------------------- training callbacks
// HoughTreeFunctions as TreeLearnFunctions (houghforest.rs:196-311).  The trainer that calls them is
// stamm's (not vendored: UNPINNED); these are the depthhead-owned pieces it calls.

// binarize (houghforest.rs:185-193) of a candidate NodeParam on training samples: patches
// [n][sh][sw] u16 (InMutSubImage crops, prediction.rs:223-226), rect = x, y, w, h per rectangle.
void orc_train_binarize(const uint16_t* patches, uint32_t sw, uint32_t sh, const uint32_t* idx, uint64_t n_idx,
                        const int32_t* rects /*8: r1 x0,y0,x1,y1, r2 ..*/, double threshold, uint8_t* bits) {
    Rect sub = rect_new(0, 0, sw, sh);
    Rect r1 = rect_new((uint32_t)rects[0], (uint32_t)rects[1], (uint32_t)(rects[2] - rects[0]), (uint32_t)(rects[3] - rects[1]));
    Rect r2 = rect_new((uint32_t)rects[4], (uint32_t)rects[5], (uint32_t)(rects[6] - rects[4]), (uint32_t)(rects[7] - rects[5]));
    for (uint64_t k = 0; k < n_idx; ++k) {
        Image img{patches + (size_t)idx[k] * sw * sh, sw, sh};
        const double a1 = average_value_in_rect(img, sub, r1), a2 = average_value_in_rect(img, sub, r2);
        bits[k] = (a1 - a2 > threshold) ? 1 : 0;
    }
}

struct OrcSideStats {  // what the GPU scorer reports per (candidate, side)
    uint64_t n, n_pos;
    double det_off, det_rot;  // determinants of the two covariance matrices (NaN for fewer than two positives)
};

static inline double rs_ln(double x) { return x == 0.0 ? 0.0 : std::log(x); }  // ln! (houghforest.rs:18-20)

// entropy (houghforest.rs:258-264)
static double train_entropy(const uint8_t* is_obj, const uint32_t* idx, uint64_t n) {
    uint64_t positives = 0;
    for (uint64_t k = 0; k < n; ++k) positives += is_obj[idx[k]] ? 1 : 0;
    const double prob = (double)positives / (double)n;  // rel!
    return prob * rs_ln(prob) + (1.0 - prob) * rs_ln(1.0 - prob);
}
// regression_log (houghforest.rs:265-284); *bad = 1 where the reference hits unreachable!()
static double train_regression_log(const uint8_t* is_obj, const float* offsets, const double* rotations, const uint32_t* idx,
                                   uint64_t n, OrcSideStats* st, int* bad) {
    std::vector<double> off, rot;
    for (uint64_t k = 0; k < n; ++k)
        if (is_obj[idx[k]]) {
            for (int c = 0; c < 3; ++c) {
                off.push_back((double)offsets[(size_t)idx[k] * 3 + c]);   // x.0[c] as f64
                rot.push_back(rotations[(size_t)idx[k] * 3 + c]);
            }
        }
    st->n = n;
    st->n_pos = off.size() / 3;
    st->det_off = st->det_rot = std::numeric_limits<double>::quiet_NaN();
    if (off.empty()) return 0.0;
    double mo[3], co[9], mr[3], cr[9];
    estimate_mean_cov<double, 3>(off.data(), off.size() / 3, mo, co);
    estimate_mean_cov<double, 3>(rot.data(), rot.size() / 3, mr, cr);
    st->det_off = mat3_det<double>(co);
    st->det_rot = mat3_det<double>(cr);
    const double x = st->det_off + st->det_rot;
    if (x > 0.0) return std::log(x);
    if (x < -0.001) { *bad = 1; return 0.0; }
    return 0.0;
}
// impurity (houghforest.rs:250-295).  Returns NaN-free values only for non-empty sides (an empty
// side makes the reference's assert!(res.is_finite()) fire: rel!(0, 0) is NaN).
double orc_train_impurity(const uint8_t* is_obj, const float* offsets, const double* rotations, const uint32_t* left,
                          uint64_t nl, const uint32_t* right, uint64_t nr, uint64_t depth, double steepness,
                          OrcSideStats* stats /*[2] or NULL*/, int* bad) {
    OrcSideStats st[2];
    int b = 0;
    const uint64_t count = nl + nr;
    const double left_factor = (double)nl / (double)count, right_factor = (double)nr / (double)count;
    const double impurity = -(left_factor * train_entropy(is_obj, left, nl) + right_factor * train_entropy(is_obj, right, nr));
    const double rl = train_regression_log(is_obj, offsets, rotations, left, nl, &st[0], &b);
    const double rr = train_regression_log(is_obj, offsets, rotations, right, nr, &st[1], &b);
    const double regression_uncert = left_factor * rl + right_factor * rr;
    const double e_factor = -((double)depth / steepness);
    const double f = std::exp(e_factor);
    if (stats) { stats[0] = st[0]; stats[1] = st[1]; }
    if (bad) *bad = b;
    return impurity + (1.0 - f) * regression_uncert;
}
// the same combination from per-side statistics (what the product does with the GPU's numbers)
double orc_train_impurity_from_stats(const OrcSideStats* st /*[2]*/, uint64_t depth, double steepness, int* bad) {
    auto entropy = [](const OrcSideStats& s) {
        const double prob = (double)s.n_pos / (double)s.n;
        return prob * rs_ln(prob) + (1.0 - prob) * rs_ln(1.0 - prob);
    };
    auto reglog = [&](const OrcSideStats& s) {
        if (s.n_pos == 0) return 0.0;
        const double x = s.det_off + s.det_rot;
        if (x > 0.0) return std::log(x);
        if (x < -0.001 && bad) *bad = 1;
        return 0.0;
    };
    const uint64_t count = st[0].n + st[1].n;
    const double lf = (double)st[0].n / (double)count, rf = (double)st[1].n / (double)count;
    const double impurity = -(lf * entropy(st[0]) + rf * entropy(st[1]));
    const double reg = lf * reglog(st[0]) + rf * reglog(st[1]);
    const double f = std::exp(-((double)depth / steepness));
    return impurity + (1.0 - f) * reg;
}
// early_stop (houghforest.rs:302-311)
int orc_train_early_stop(const uint8_t* is_obj, const uint32_t* idx, uint64_t n, uint64_t depth, uint64_t max_depth,
                         uint64_t min_subset_size) {
    bool all_neg = true;
    for (uint64_t k = 0; k < n; ++k)
        if (is_obj[idx[k]]) all_neg = false;
    if (all_neg) return 1;
    return (depth >= max_depth || n < min_subset_size) ? 1 : 0;
}

}  // extern "C"
