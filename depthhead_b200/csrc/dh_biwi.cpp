// dh_biwi.cpp — host-side readers of the Biwi Kinect Head Pose Database side files, with the
// behaviour of the reference's src/db_reader/biwi.rs: read_cal (:27-60), read_gt (:63-77) and the
// header of read_depth (:81-86).  The depth runs themselves are expanded on the GPU
// (biwi_decode_kernel in dh_kernels.cu).
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/depthhead_cuda.h"
#include "dh_forest.hpp"

namespace dh {

namespace {
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }
inline uint32_t le_u32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline float le_f32(const uint8_t* p) {
    const uint32_t u = le_u32(p);
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
}  // namespace

// read_depth's header (biwi.rs:83-84)
void biwi_depth_dims(const uint8_t* file, size_t len, uint32_t* w, uint32_t* h) {
    if (len < 8) throw ModelError(DH_E_ARG, "Biwi depth file shorter than its 8-byte header");
    *w = le_u32(file);
    *h = le_u32(file + 4);
}

// read_cal (biwi.rs:27-60).  Per line the reference collects the non-overlapping matches of the
// regex (\d+[\.\d+]*) — a run of digits followed by any mix of '.', digits and '+' — and parses
// each with f32::from_str; exactly three per line, for the first three lines.  A '-' is never part
// of a match, so signs are dropped, exactly as there.  (Rust's \d also matches non-ASCII decimal
// digits, which f32::from_str then rejects; here they are simply not digits: both end in an
// error unless the line still holds three ASCII numbers.)
void biwi_parse_cal(const char* text, size_t len, float K[9]) {
    size_t pos = 0;
    for (int j = 0; j < 3; ++j) {
        // read_line: up to and including '\n' (an absent line reads as empty)
        size_t end = pos;
        while (end < len && text[end] != '\n') ++end;
        const size_t line_end = end;
        int found = 0;
        size_t i = pos;
        while (i < line_end) {
            if (!is_digit(text[i])) {
                ++i;
                continue;
            }
            size_t k = i;
            while (k < line_end && is_digit(text[k])) ++k;
            while (k < line_end && (is_digit(text[k]) || text[k] == '.' || text[k] == '+')) ++k;
            if (found == 3) throw ModelError(DH_E_ARG, "depth.cal: more than three numbers on a line (the reference indexes out of bounds, biwi.rs:45)");
            // f32::from_str on the token: digits, optionally '.' and more digits; anything else
            // the character class lets through ("1.2.3", "4+5") is a ParseFloatError
            const std::string tok(text + i, k - i);
            size_t d = 0;
            while (d < tok.size() && is_digit(tok[d])) ++d;
            if (d < tok.size() && tok[d] == '.') {
                ++d;
                while (d < tok.size() && is_digit(tok[d])) ++d;
            }
            if (d != tok.size()) throw ModelError(DH_E_ARG, "depth.cal: `" + tok + "` is not a number (ParseFloatError in the reference)");
            K[j * 3 + found] = std::strtof(tok.c_str(), nullptr);  // correctly rounded, like Rust
            ++found;
            i = k;
        }
        if (found != 3) throw ModelError(DH_E_ARG, "depth.cal: Unsupported Calibration-File (a line without exactly three numbers)");
        pos = end < len ? end + 1 : len;
    }
}

// read_gt (biwi.rs:63-77) + IntrinsicMatrix::space_to_img_coord (types.rs:424-428)
void biwi_parse_pose(const uint8_t* file, size_t len, const float K[9], float pos3d[3], float pos2d[2], float rot[3]) {
    if (len < 24) throw ModelError(DH_E_ARG, "Biwi pose file shorter than six f32 (UnexpectedEof in the reference)");
    float r[6];
    for (int i = 0; i < 6; ++i) r[i] = le_f32(file + 4 * i);
    for (int i = 0; i < 3; ++i) {
        pos3d[i] = r[i];
        rot[i] = r[3 + i];
    }
    float q[3];
    for (int j = 0; j < 3; ++j) {  // Mat3 * Vec3, meancov_estimation.rs:201-216: separate multiplications and additions
        volatile float t = r[0] * K[j * 3 + 0];
        volatile float u = r[1] * K[j * 3 + 1];
        t = t + u;
        u = r[2] * K[j * 3 + 2];
        t = t + u;
        q[j] = t;
    }
    pos2d[0] = q[0] / q[2];
    pos2d[1] = q[1] / q[2];
}

}  // namespace dh
