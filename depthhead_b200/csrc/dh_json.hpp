// dh_json.hpp — minimal strict JSON pull parser for the HoughPrediction document.
//
// The reference loads its model with serde_json::from_str::<HoughPrediction> (Readme.md:82-86).
// A trained forest is a 100 MB+ document, so this is a single-pass, allocation-free reader with a
// schema-driven consumer (dh_forest.cpp) rather than a DOM.  Number semantics follow serde_json:
// integer fields accept only integer literals; f64 fields accept any number and are converted
// with a correctly-rounded parser (std::from_chars); f32 fields are parsed as f64 and narrowed.
#pragma once

#include <charconv>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <string_view>

namespace dh {

struct JsonError : std::runtime_error {
    size_t offset;
    JsonError(const std::string& m, size_t off)
        : std::runtime_error(m + " at byte " + std::to_string(off)), offset(off) {}
};

class JsonReader {
public:
    JsonReader(const char* s, size_t n) : b_(s), p_(s), e_(s + n) {}

    size_t offset() const { return (size_t)(p_ - b_); }
    [[noreturn]] void fail(const std::string& m) const { throw JsonError(m, offset()); }

    void ws() {
        while (p_ < e_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\r' || *p_ == '\t')) ++p_;
    }
    char peek() {
        ws();
        if (p_ >= e_) fail("unexpected end of document");
        return *p_;
    }
    void expect(char c) {
        if (peek() != c) fail(std::string("expected '") + c + "'");
        ++p_;
    }
    bool consume(char c) {
        if (peek() == c) {
            ++p_;
            return true;
        }
        return false;
    }
    void expect_end() {
        ws();
        if (p_ != e_) fail("trailing characters after the document");
    }

    // Strings: keys in this schema are plain ASCII; escapes are decoded anyway.
    std::string parse_string() {
        expect('"');
        std::string out;
        while (true) {
            if (p_ >= e_) fail("unterminated string");
            unsigned char c = (unsigned char)*p_++;
            if (c == '"') break;
            if (c < 0x20) fail("control character in string");
            if (c != '\\') {
                out.push_back((char)c);
                continue;
            }
            if (p_ >= e_) fail("unterminated escape");
            char esc = *p_++;
            switch (esc) {
                case '"': out.push_back('"'); break;
                case '\\': out.push_back('\\'); break;
                case '/': out.push_back('/'); break;
                case 'b': out.push_back('\b'); break;
                case 'f': out.push_back('\f'); break;
                case 'n': out.push_back('\n'); break;
                case 'r': out.push_back('\r'); break;
                case 't': out.push_back('\t'); break;
                case 'u': {
                    if (e_ - p_ < 4) fail("short \\u escape");
                    unsigned v = 0;
                    for (int i = 0; i < 4; ++i) {
                        char h = *p_++;
                        v <<= 4;
                        if (h >= '0' && h <= '9') v |= (unsigned)(h - '0');
                        else if (h >= 'a' && h <= 'f') v |= (unsigned)(h - 'a' + 10);
                        else if (h >= 'A' && h <= 'F') v |= (unsigned)(h - 'A' + 10);
                        else fail("bad \\u escape");
                    }
                    // UTF-8 encode (surrogate pairs are not needed for this schema's keys)
                    if (v < 0x80) out.push_back((char)v);
                    else if (v < 0x800) {
                        out.push_back((char)(0xC0 | (v >> 6)));
                        out.push_back((char)(0x80 | (v & 0x3F)));
                    } else {
                        out.push_back((char)(0xE0 | (v >> 12)));
                        out.push_back((char)(0x80 | ((v >> 6) & 0x3F)));
                        out.push_back((char)(0x80 | (v & 0x3F)));
                    }
                    break;
                }
                default: fail("bad escape");
            }
        }
        return out;
    }

    // Scans one JSON number token and returns [begin,end); is_int = no fraction / exponent.
    std::string_view number_token(bool* is_int) {
        ws();
        const char* s = p_;
        if (p_ < e_ && *p_ == '-') ++p_;
        if (p_ >= e_ || *p_ < '0' || *p_ > '9') fail("expected a number");
        if (*p_ == '0') ++p_;
        else while (p_ < e_ && *p_ >= '0' && *p_ <= '9') ++p_;
        bool integer = true;
        if (p_ < e_ && *p_ == '.') {
            integer = false;
            ++p_;
            if (p_ >= e_ || *p_ < '0' || *p_ > '9') fail("digit expected after '.'");
            while (p_ < e_ && *p_ >= '0' && *p_ <= '9') ++p_;
        }
        if (p_ < e_ && (*p_ == 'e' || *p_ == 'E')) {
            integer = false;
            ++p_;
            if (p_ < e_ && (*p_ == '+' || *p_ == '-')) ++p_;
            if (p_ >= e_ || *p_ < '0' || *p_ > '9') fail("digit expected in exponent");
            while (p_ < e_ && *p_ >= '0' && *p_ <= '9') ++p_;
        }
        if (is_int) *is_int = integer;
        return std::string_view(s, (size_t)(p_ - s));
    }

    double parse_f64() {
        bool is_int;
        std::string_view t = number_token(&is_int);
        double v = 0.0;
        auto r = std::from_chars(t.data(), t.data() + t.size(), v);
        if (r.ec == std::errc::result_out_of_range) {
            // from_chars leaves v untouched here: overflow is an error (serde_json: "number out
            // of range"), underflow rounds toward zero / a subnormal.
            std::string tmp(t);
            v = std::strtod(tmp.c_str(), nullptr);
            if (v > 1.0e308 || v < -1.0e308) fail("number out of range");
        } else if (r.ec != std::errc() || r.ptr != t.data() + t.size()) {
            fail("malformed number");
        }
        return v;
    }
    float parse_f32() { return (float)parse_f64(); }  // serde: visit_f64 then `as f32`

    uint64_t parse_u64() {
        bool is_int;
        size_t off = offset();
        std::string_view t = number_token(&is_int);
        if (!is_int || t[0] == '-') throw JsonError("expected an unsigned integer", off);
        uint64_t v = 0;
        auto r = std::from_chars(t.data(), t.data() + t.size(), v);
        if (r.ec != std::errc() || r.ptr != t.data() + t.size()) throw JsonError("integer out of range", off);
        return v;
    }
    int64_t parse_i64() {
        bool is_int;
        size_t off = offset();
        std::string_view t = number_token(&is_int);
        if (!is_int) throw JsonError("expected an integer", off);
        int64_t v = 0;
        auto r = std::from_chars(t.data(), t.data() + t.size(), v);
        if (r.ec != std::errc() || r.ptr != t.data() + t.size()) throw JsonError("integer out of range", off);
        return v;
    }
    uint32_t parse_u32() {
        size_t off = offset();
        uint64_t v = parse_u64();
        if (v > 0xFFFFFFFFull) throw JsonError("integer does not fit u32", off);
        return (uint32_t)v;
    }

    void skip_literal(const char* lit) {
        size_t n = std::strlen(lit);
        if ((size_t)(e_ - p_) < n || std::memcmp(p_, lit, n) != 0) fail("bad literal");
        p_ += n;
    }
    void skip_value() {
        char c = peek();
        if (c == '{') {
            ++p_;
            if (consume('}')) return;
            do {
                parse_string();
                expect(':');
                skip_value();
            } while (consume(','));
            expect('}');
        } else if (c == '[') {
            ++p_;
            if (consume(']')) return;
            do skip_value(); while (consume(','));
            expect(']');
        } else if (c == '"') {
            parse_string();
        } else if (c == 't') skip_literal("true");
        else if (c == 'f') skip_literal("false");
        else if (c == 'n') skip_literal("null");
        else number_token(nullptr);
    }

    // on_key(key) must consume exactly one value.
    template <class F>
    void parse_object(F&& on_key) {
        expect('{');
        if (consume('}')) return;
        do {
            std::string key = parse_string();
            expect(':');
            on_key(key);
        } while (consume(','));
        expect('}');
    }
    // on_elem(index) must consume exactly one value.
    template <class F>
    void parse_array(F&& on_elem) {
        expect('[');
        if (consume(']')) return;
        size_t i = 0;
        do on_elem(i++); while (consume(','));
        expect(']');
    }

private:
    const char* b_;
    const char* p_;
    const char* e_;
};

}  // namespace dh
