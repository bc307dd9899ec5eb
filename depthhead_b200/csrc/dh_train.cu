// dh_train.cu — split scoring for Hough-forest TRAINING.
//
// The reference trains with stamm's `train_forest_parallel` (a crate that is not vendored), which
// calls back into depthhead's HoughTreeFunctions (src/hough/houghforest.rs:196-311).  What costs
// the time there is, per tree node, `impurity` (:250-295) of every candidate NodeParam: binarize
// (:185-193) every sample of the node, then entropy and the covariance determinants of the
// offsets and rotations on either side.  These kernels do that for all nodes of one tree level
// at once; the tree-growing loop above them is host code (depthhead_b200/train.py).
//
// Exactness: the covariance sums run over the samples IN SET ORDER with separate multiplications
// and additions, exactly estimate_mean_cov (meancov_estimation.rs:359-378) and Mat3::det
// (:339-343) in f64, so the per-side statistics equal the CPU restatement bit for bit; ln/exp of
// the final combination are taken on the host with the same libm the reference's f64::ln uses.
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "dh_ctx.hpp"

namespace dh {

#define DH_CUDA_T(expr)                                                                                 \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess)                                                                          \
            throw ModelError(DH_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));            \
    } while (0)

struct TrainSet {
    int device = 0;
    uint64_t n = 0;
    uint32_t sw = 0, sh = 0, rw = 0, rh = 0, bw = 0, bh = 0, bpitch = 0;
    uint32_t* box = nullptr;      // [n][bh][bpitch] box sums of every sample patch
    uint8_t* is_obj = nullptr;    // [n]
    double* off = nullptr;        // [n][3] offsets widened to f64 (`x.0[k] as f64`, houghforest.rs:268-269)
    double* rot = nullptr;        // [n][3]
    ~TrainSet() {
        cudaSetDevice(device);
        if (box) cudaFree(box);
        if (is_obj) cudaFree(is_obj);
        if (off) cudaFree(off);
        if (rot) cudaFree(rot);
    }
};

namespace dev {

struct Cand {  // one candidate NodeParam: box-table offsets of its two rectangles + threshold
    uint32_t tap1, tap2;
    double threshold;
};

// binarize (houghforest.rs:185-193) from two box sums: avg = sum as f64 / count as f64
__device__ __forceinline__ bool train_bit(const uint32_t* __restrict__ tbl, const Cand& c, double count) {
    const double a1 = __ddiv_rn((double)__ldg(tbl + c.tap1), count), a2 = __ddiv_rn((double)__ldg(tbl + c.tap2), count);
    return __dsub_rn(a1, a2) > c.threshold;
}

// One thread per (node, candidate).  Pass 1: side sizes, positives and the sums for the two
// means; pass 2: the two covariance matrices per side; then the four determinants.
__global__ void __launch_bounds__(128) train_score_kernel(const uint32_t* __restrict__ box, uint32_t tbl_words,
                                                          const uint8_t* __restrict__ is_obj, const double* __restrict__ off,
                                                          const double* __restrict__ rot, const uint32_t* __restrict__ idx,
                                                          const unsigned long long* __restrict__ node_off,
                                                          const Cand* __restrict__ cands, uint32_t m, double count,
                                                          dh_split_stats* __restrict__ out) {
    const uint32_t node = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= m) return;
    const Cand cd = cands[(size_t)node * m + c];
    const unsigned long long s0 = node_off[node], s1 = node_off[node + 1];
    uint32_t n[2] = {0, 0}, np[2] = {0, 0};
    double mo[2][3], mr[2][3];
    for (unsigned long long s = s0; s < s1; ++s) {
        const uint32_t j = __ldg(idx + s);
        const int side = train_bit(box + (size_t)j * tbl_words, cd, count) ? 1 : 0;
        ++n[side];
        if (__ldg(is_obj + j)) {
            const double* o = off + (size_t)j * 3;
            const double* r = rot + (size_t)j * 3;
            if (np[side] == 0) {  // mean starts from the first element (meancov_estimation.rs:363-366)
                for (int k = 0; k < 3; ++k) { mo[side][k] = __ldg(o + k); mr[side][k] = __ldg(r + k); }
            } else {
                for (int k = 0; k < 3; ++k) { mo[side][k] = __dadd_rn(mo[side][k], __ldg(o + k)); mr[side][k] = __dadd_rn(mr[side][k], __ldg(r + k)); }
            }
            ++np[side];
        }
    }
    for (int sd = 0; sd < 2; ++sd)
        if (np[sd])
            for (int k = 0; k < 3; ++k) {
                mo[sd][k] = __ddiv_rn(mo[sd][k], (double)np[sd]);
                mr[sd][k] = __ddiv_rn(mr[sd][k], (double)np[sd]);
            }
    // covariance: upper triangle (d[a]*d[b] == d[b]*d[a] exactly), entries 00 01 02 11 12 22
    double co[2][6], cr[2][6];
    uint32_t seen[2] = {0, 0};
    for (unsigned long long s = s0; s < s1; ++s) {
        const uint32_t j = __ldg(idx + s);
        if (!__ldg(is_obj + j)) continue;
        const int side = train_bit(box + (size_t)j * tbl_words, cd, count) ? 1 : 0;
        double d[3], e[3];
        for (int k = 0; k < 3; ++k) {
            d[k] = __dsub_rn(__ldg(off + (size_t)j * 3 + k), mo[side][k]);
            e[k] = __dsub_rn(__ldg(rot + (size_t)j * 3 + k), mr[side][k]);
        }
        const double po[6] = {__dmul_rn(d[0], d[0]), __dmul_rn(d[0], d[1]), __dmul_rn(d[0], d[2]),
                              __dmul_rn(d[1], d[1]), __dmul_rn(d[1], d[2]), __dmul_rn(d[2], d[2])};
        const double pr[6] = {__dmul_rn(e[0], e[0]), __dmul_rn(e[0], e[1]), __dmul_rn(e[0], e[2]),
                              __dmul_rn(e[1], e[1]), __dmul_rn(e[1], e[2]), __dmul_rn(e[2], e[2])};
        if (seen[side] == 0) {
            for (int k = 0; k < 6; ++k) { co[side][k] = po[k]; cr[side][k] = pr[k]; }
        } else {
            for (int k = 0; k < 6; ++k) { co[side][k] = __dadd_rn(co[side][k], po[k]); cr[side][k] = __dadd_rn(cr[side][k], pr[k]); }
        }
        ++seen[side];
    }
    dh_split_stats st;
    for (int sd = 0; sd < 2; ++sd) {
        st.n[sd] = n[sd];
        st.n_pos[sd] = np[sd];
        double det_o = __longlong_as_double(0x7ff8000000000000ll), det_r = det_o;
        if (np[sd]) {
            const double dn1 = (double)(np[sd] - 1u);  // n == 1: 0/0 = NaN entries, NaN determinant
            double a[6], b[6];
            for (int k = 0; k < 6; ++k) { a[k] = __ddiv_rn(co[sd][k], dn1); b[k] = __ddiv_rn(cr[sd][k], dn1); }
            // Mat3::det (meancov_estimation.rs:339-343) with m10 = m01, m20 = m02, m21 = m12:
            // m00*(m11*m22 - m12*m21) - m10*(m01*m22 - m02*m21) + m20*(m01*m12 - m02*m11)
            auto det = [](const double* q) {
                const double t0 = __dmul_rn(q[0], __dsub_rn(__dmul_rn(q[3], q[5]), __dmul_rn(q[4], q[4])));
                const double t1 = __dmul_rn(q[1], __dsub_rn(__dmul_rn(q[1], q[5]), __dmul_rn(q[2], q[4])));
                const double t2 = __dmul_rn(q[2], __dsub_rn(__dmul_rn(q[1], q[4]), __dmul_rn(q[2], q[3])));
                return __dadd_rn(__dsub_rn(t0, t1), t2);
            };
            det_o = det(a);
            det_r = det(b);
        }
        st.det_off[sd] = det_o;
        st.det_rot[sd] = det_r;
    }
    st.impurity = 0.0;
    out[(size_t)node * m + c] = st;
}

// bits of ONE chosen candidate per node over the node's samples
__global__ void __launch_bounds__(256) train_split_kernel(const uint32_t* __restrict__ box, uint32_t tbl_words,
                                                          const uint32_t* __restrict__ idx,
                                                          const unsigned long long* __restrict__ node_off,
                                                          const Cand* __restrict__ chosen, uint32_t n_nodes, double count,
                                                          uint8_t* __restrict__ bits) {
    const unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= node_off[n_nodes]) return;
    uint32_t lo = 0, hi = n_nodes;  // node of sample slot s: last k with node_off[k] <= s
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (node_off[mid] <= s) lo = mid; else hi = mid;
    }
    bits[s] = train_bit(box + (size_t)__ldg(idx + s) * tbl_words, chosen[lo], count) ? 1 : 0;
}

}  // namespace dev

// ---------------------------------------------------------------------------------------------- host
namespace {
template <typename T>
struct DevBuf {
    T* p = nullptr;
    explicit DevBuf(size_t n) { if (cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) { p = nullptr; throw ModelError(DH_E_CUDA, "cudaMalloc failed (training scratch)"); } }
    ~DevBuf() { if (p) cudaFree(p); }
};

// `ln!` (houghforest.rs:18-20)
inline double rs_ln(double x) { return x == 0.0 ? 0.0 : std::log(x); }

// impurity (houghforest.rs:250-295) from the per-side statistics; NaN where the reference itself
// aborts: an empty side (rel!(0, 0) is NaN and assert!(res.is_finite()) fires, :293) or a
// covariance determinant sum below -0.001 (unreachable!(), :281 — rounding of a rank-deficient
// covariance of widely spread samples can get there)
double impurity_from_stats(const dh_split_stats& s, uint64_t depth, double steepness, bool* unreachable) {
    if (s.n[0] == 0 || s.n[1] == 0) return std::numeric_limits<double>::quiet_NaN();
    auto entropy = [](uint32_t pos, uint32_t n) {
        const double prob = (double)pos / (double)n;
        return prob * rs_ln(prob) + (1.0 - prob) * rs_ln(1.0 - prob);
    };
    auto reglog = [&](int sd) {
        if (s.n_pos[sd] == 0) return 0.0;
        const double x = s.det_off[sd] + s.det_rot[sd];
        if (x > 0.0) return std::log(x);
        if (x < -0.001) *unreachable = true;  // houghforest.rs:281
        return 0.0;
    };
    const uint64_t count = (uint64_t)s.n[0] + s.n[1];
    const double lf = (double)s.n[0] / (double)count, rf = (double)s.n[1] / (double)count;
    const double impurity = -(lf * entropy(s.n_pos[0], s.n[0]) + rf * entropy(s.n_pos[1], s.n[1]));
    const double reg = lf * reglog(0) + rf * reglog(1);
    const double f = std::exp(-((double)depth / steepness));
    return impurity + (1.0 - f) * reg;
}

void make_cands(const TrainSet& ts, const int32_t* rects, const double* thr, size_t n, std::vector<dev::Cand>& out) {
    out.resize(n);
    for (size_t i = 0; i < n; ++i) {
        const int32_t* r = rects + i * 8;
        for (int k = 0; k < 2; ++k) {
            const int32_t x0 = r[k * 4], y0 = r[k * 4 + 1], x1 = r[k * 4 + 2], y1 = r[k * 4 + 3];
            if (x0 < 0 || y0 < 0 || x1 - x0 != (int32_t)ts.rw || y1 - y0 != (int32_t)ts.rh || x1 > (int32_t)ts.sw || y1 > (int32_t)ts.sh)
                throw ModelError(DH_E_ARG, "candidate " + std::to_string(i) + ": rectangle is not " + std::to_string(ts.rw) + "x" +
                                               std::to_string(ts.rh) + " inside the sub-image (this training set was built for one rectangle size)");
            (k ? out[i].tap2 : out[i].tap1) = (uint32_t)y0 * ts.bpitch + (uint32_t)x0;
        }
        out[i].threshold = thr[i];
    }
}
}  // namespace

TrainSet* Context::trainset_create(const uint16_t* patches, uint64_t n, uint32_t sw, uint32_t sh, uint32_t rw, uint32_t rh,
                                   const uint8_t* is_object, const float* offsets, const double* rotations) {
    DH_CUDA_T(cudaSetDevice(device_));
    if (n == 0 || n > 0x7fffffffull) throw ModelError(DH_E_ARG, "training set must hold 1 .. 2^31-1 samples");
    if (!box_image_supported(sw, sh, sw, sh, rw, rh) || (uint64_t)rw * rh > 16383u)
        throw ModelError(DH_E_SHAPE, "feature rectangle size not supported by the box-sum tables (see DESIGN.md)");
    std::unique_ptr<TrainSet> ts(new TrainSet());
    ts->device = device_;
    ts->n = n; ts->sw = sw; ts->sh = sh; ts->rw = rw; ts->rh = rh;
    ts->bw = sw - rw + 1; ts->bh = sh - rh + 1; ts->bpitch = (ts->bw + 3) & ~3u;
    DH_CUDA_T(cudaMalloc((void**)&ts->box, n * (size_t)ts->bh * ts->bpitch * sizeof(uint32_t)));
    DH_CUDA_T(cudaMalloc((void**)&ts->is_obj, n));
    DH_CUDA_T(cudaMalloc((void**)&ts->off, n * 3 * sizeof(double)));
    DH_CUDA_T(cudaMalloc((void**)&ts->rot, n * 3 * sizeof(double)));
    std::vector<double> off64(n * 3);
    for (size_t i = 0; i < n * 3; ++i) off64[i] = (double)offsets[i];
    DH_CUDA_T(cudaMemcpyAsync(ts->is_obj, is_object, n, cudaMemcpyHostToDevice, stream_));
    DH_CUDA_T(cudaMemcpyAsync(ts->off, off64.data(), n * 3 * sizeof(double), cudaMemcpyHostToDevice, stream_));
    DH_CUDA_T(cudaMemcpyAsync(ts->rot, rotations, n * 3 * sizeof(double), cudaMemcpyHostToDevice, stream_));
    // box-sum table of every patch: the prediction path's front-end kernel with frame = patch
    const size_t patch_px = (size_t)sw * sh;
    const uint64_t chunk = std::max<uint64_t>(1, std::min<uint64_t>(n, (256ull << 20) / (patch_px * 2)));
    DevBuf<uint16_t> d_px(chunk * patch_px);
    Geometry g{};
    g.w = sw; g.h = sh; g.sw = sw; g.sh = sh; g.rw = rw; g.rh = rh;
    g.box_w = ts->bw; g.box_h = ts->bh; g.box_pitch = ts->bpitch;
    for (uint64_t f0 = 0; f0 < n; f0 += chunk) {
        const uint64_t nc = std::min(chunk, n - f0);
        DH_CUDA_T(cudaMemcpyAsync(d_px.p, patches + f0 * patch_px, nc * patch_px * sizeof(uint16_t), cudaMemcpyHostToDevice, stream_));
        FrameBuffers b{};
        b.depth = d_px.p;
        b.box = ts->box + f0 * (size_t)ts->bh * ts->bpitch;
        launch_box_image(b, g, (uint32_t)nc, n_sms_, stream_);
        DH_CUDA_T(cudaGetLastError());
        DH_CUDA_T(cudaStreamSynchronize(stream_));
    }
    return ts.release();
}

void Context::train_score_level(const TrainSet& ts, const uint32_t* sample_idx, const uint64_t* node_off, uint32_t n_nodes,
                                const int32_t* cand_rects, const double* cand_thr, uint32_t m, uint64_t depth, double steepness,
                                dh_split_stats* out) {
    DH_CUDA_T(cudaSetDevice(device_));
    if (n_nodes == 0 || m == 0) return;
    const uint64_t total = node_off[n_nodes];
    for (uint32_t k = 0; k < n_nodes; ++k)
        if (node_off[k + 1] < node_off[k]) throw ModelError(DH_E_ARG, "node_off is not monotone");
    for (uint64_t s = 0; s < total; ++s)
        if (sample_idx[s] >= ts.n) throw ModelError(DH_E_ARG, "sample index out of range");
    std::vector<dev::Cand> cands;
    make_cands(ts, cand_rects, cand_thr, (size_t)n_nodes * m, cands);
    DevBuf<uint32_t> d_idx(total);
    DevBuf<unsigned long long> d_off(n_nodes + 1);
    DevBuf<dev::Cand> d_c(cands.size());
    DevBuf<dh_split_stats> d_out(cands.size());
    DH_CUDA_T(cudaMemcpyAsync(d_idx.p, sample_idx, total * sizeof(uint32_t), cudaMemcpyHostToDevice, stream_));
    DH_CUDA_T(cudaMemcpyAsync(d_off.p, node_off, (n_nodes + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, stream_));
    DH_CUDA_T(cudaMemcpyAsync(d_c.p, cands.data(), cands.size() * sizeof(dev::Cand), cudaMemcpyHostToDevice, stream_));
    dim3 gr((m + 127) / 128, n_nodes);
    dev::train_score_kernel<<<gr, 128, 0, stream_>>>(ts.box, ts.bh * ts.bpitch, ts.is_obj, ts.off, ts.rot, d_idx.p, d_off.p, d_c.p, m,
                                                     (double)((uint64_t)ts.rw * ts.rh), d_out.p);
    DH_CUDA_T(cudaGetLastError());
    DH_CUDA_T(cudaMemcpyAsync(out, d_out.p, cands.size() * sizeof(dh_split_stats), cudaMemcpyDeviceToHost, stream_));
    DH_CUDA_T(cudaStreamSynchronize(stream_));
    for (size_t i = 0; i < cands.size(); ++i) {
        bool unreachable = false;
        out[i].impurity = impurity_from_stats(out[i], depth, steepness, &unreachable);
        if (unreachable) out[i].impurity = std::numeric_limits<double>::quiet_NaN();
    }
}

void Context::train_split_level(const TrainSet& ts, const uint32_t* sample_idx, const uint64_t* node_off, uint32_t n_nodes,
                                const int32_t* rects, const double* thr, uint8_t* bits) {
    DH_CUDA_T(cudaSetDevice(device_));
    if (n_nodes == 0) return;
    const uint64_t total = node_off[n_nodes];
    if (total == 0) return;
    for (uint64_t s = 0; s < total; ++s)
        if (sample_idx[s] >= ts.n) throw ModelError(DH_E_ARG, "sample index out of range");
    std::vector<dev::Cand> cands;
    make_cands(ts, rects, thr, n_nodes, cands);
    DevBuf<uint32_t> d_idx(total);
    DevBuf<unsigned long long> d_off(n_nodes + 1);
    DevBuf<dev::Cand> d_c(cands.size());
    DevBuf<uint8_t> d_bits(total);
    DH_CUDA_T(cudaMemcpyAsync(d_idx.p, sample_idx, total * sizeof(uint32_t), cudaMemcpyHostToDevice, stream_));
    DH_CUDA_T(cudaMemcpyAsync(d_off.p, node_off, (n_nodes + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, stream_));
    DH_CUDA_T(cudaMemcpyAsync(d_c.p, cands.data(), cands.size() * sizeof(dev::Cand), cudaMemcpyHostToDevice, stream_));
    dev::train_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream_>>>(ts.box, ts.bh * ts.bpitch, d_idx.p, d_off.p, d_c.p, n_nodes,
                                                                                 (double)((uint64_t)ts.rw * ts.rh), d_bits.p);
    DH_CUDA_T(cudaGetLastError());
    DH_CUDA_T(cudaMemcpyAsync(bits, d_bits.p, total, cudaMemcpyDeviceToHost, stream_));
    DH_CUDA_T(cudaStreamSynchronize(stream_));
}

void trainset_free(TrainSet* t) { delete t; }

}  // namespace dh
