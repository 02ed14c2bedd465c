// ingest.cpp -- multi-threaded reader of HyPhy site-rate JSON files into one CSR pack (host only).
//
// Replaces the read half of parse_site_rates (tapir/compute.py:30-39) for MANY files at once:
//   rates = [line["rate"] for line in data["sites"]["rates"]] / correction
// The files are the output of tapir/data/models_and_rates.bf:1018-1104, possibly already
// rewritten by the reference with a "corrected_rates" list (ignored here, as the reference ignores it).
// A small strict JSON walker is used rather than pattern matching, so key order, whitespace, nesting and
// extra keys do not matter; Python's non-standard NaN / Infinity / -Infinity tokens are accepted because
// json.load accepts them.  Numbers are converted with std::from_chars (correctly rounded, locale-free),
// which is what Python's float() yields, and the division by `correction` is the same IEEE division numpy does.
#include "../../include/tapir_cuda.h"

#include <atomic>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Cursor {
    const char* p;
    const char* end;
    std::string err;
    bool fail(const char* what)
    {
        if (err.empty()) err = what;
        return false;
    }
    void ws()
    {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
    }
    bool lit(const char* s)
    {
        const size_t n = strlen(s);
        if (size_t(end - p) >= n && memcmp(p, s, n) == 0) {
            p += n;
            return true;
        }
        return false;
    }
    // string -> out (escapes kept verbatim except that \" does not end the string; keys we look for are plain)
    bool str(std::string* out)
    {
        if (p >= end || *p != '"') return fail("expected a string");
        ++p;
        const char* s = p;
        while (p < end && *p != '"') {
            if (*p == '\\') ++p;
            ++p;
        }
        if (p >= end) return fail("unterminated string");
        if (out) out->assign(s, p - s);
        ++p;
        return true;
    }
    bool num(double* out)
    {
        ws();
        if (lit("NaN")) { *out = std::numeric_limits<double>::quiet_NaN(); return true; }
        if (lit("Infinity")) { *out = std::numeric_limits<double>::infinity(); return true; }
        if (lit("-Infinity")) { *out = -std::numeric_limits<double>::infinity(); return true; }
        const char* s = p;
        if (s < end && *s == '+') ++s;
        auto r = std::from_chars(s, end, *out);
        if (r.ec == std::errc::result_out_of_range) {  // json.load gives +-inf / 0.0 here as strtod does
            const char* q = s;
            const bool neg = q < end && *q == '-';
            while (q < end && (*q == '-' || *q == '+' || *q == '.' || *q == 'e' || *q == 'E' || (*q >= '0' && *q <= '9'))) ++q;
            bool tiny = false;
            for (const char* e = s; e < q; ++e)
                if ((*e == 'e' || *e == 'E') && e + 1 < q && e[1] == '-') tiny = true;
            *out = tiny ? (neg ? -0.0 : 0.0) : (neg ? -HUGE_VAL : HUGE_VAL);
            p = q;
            return true;
        }
        if (r.ec != std::errc()) return fail("expected a number");
        p = r.ptr;
        return true;
    }
    bool skip()  // any JSON value
    {
        ws();
        if (p >= end) return fail("unexpected end of file");
        const char c = *p;
        if (c == '"') return str(nullptr);
        if (c == '{') {
            ++p;
            ws();
            if (p < end && *p == '}') { ++p; return true; }
            for (;;) {
                ws();
                if (!str(nullptr)) return false;
                ws();
                if (p >= end || *p != ':') return fail("expected ':'");
                ++p;
                if (!skip()) return false;
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == '}') { ++p; return true; }
                return fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            ++p;
            ws();
            if (p < end && *p == ']') { ++p; return true; }
            for (;;) {
                if (!skip()) return false;
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == ']') { ++p; return true; }
                return fail("expected ',' or ']'");
            }
        }
        if (lit("true") || lit("false") || lit("null")) return true;
        double d;
        return num(&d);
    }
    // enter an object and position the cursor on the value of `key`; false (no error) if the key is absent
    bool find_key(const char* key, bool* found)
    {
        *found = false;
        ws();
        if (p >= end || *p != '{') return fail("expected an object");
        ++p;
        ws();
        if (p < end && *p == '}') { ++p; return true; }
        std::string k;
        for (;;) {
            ws();
            if (!str(&k)) return false;
            ws();
            if (p >= end || *p != ':') return fail("expected ':'");
            ++p;
            if (k == key) { *found = true; return true; }
            if (!skip()) return false;
            ws();
            if (p < end && *p == ',') { ++p; continue; }
            if (p < end && *p == '}') { ++p; return true; }
            return fail("expected ',' or '}'");
        }
    }
};

// data["sites"]["rates"][i]["rate"] for every i, appended to out
bool parse_rates(const std::string& text, std::vector<double>& out, std::string& err)
{
    Cursor c{text.data(), text.data() + text.size(), {}};
    bool found = false;
    if (!c.find_key("sites", &found)) { err = c.err; return false; }
    if (!found) { err = "no \"sites\" key"; return false; }
    if (!c.find_key("rates", &found)) { err = c.err; return false; }
    if (!found) { err = "no \"rates\" key in \"sites\""; return false; }
    c.ws();
    if (c.p >= c.end || *c.p != '[') { err = "\"rates\" is not a list"; return false; }
    ++c.p;
    c.ws();
    if (c.p < c.end && *c.p == ']') return true;
    for (;;) {
        // one {"site":..,"subst":..,"rate":..,"ll":..} record; the record must be consumed to its end
        c.ws();
        if (c.p >= c.end || *c.p != '{') { err = "a rate record is not an object"; return false; }
        ++c.p;
        bool have = false;
        double rate = 0.0;
        std::string k;
        c.ws();
        if (c.p < c.end && *c.p == '}') {
            ++c.p;
        } else {
            for (;;) {
                c.ws();
                if (!c.str(&k)) { err = c.err; return false; }
                c.ws();
                if (c.p >= c.end || *c.p != ':') { err = "expected ':'"; return false; }
                ++c.p;
                if (k == "rate") {
                    if (!c.num(&rate)) { err = c.err; return false; }
                    have = true;  // a repeated key keeps the last value, as a Python dict does
                } else if (!c.skip()) {
                    err = c.err;
                    return false;
                }
                c.ws();
                if (c.p < c.end && *c.p == ',') { ++c.p; continue; }
                if (c.p < c.end && *c.p == '}') { ++c.p; break; }
                err = "expected ',' or '}'";
                return false;
            }
        }
        if (!have) { err = "a rate record has no \"rate\" key"; return false; }  // KeyError in the reference
        out.push_back(rate);
        c.ws();
        if (c.p < c.end && *c.p == ',') { ++c.p; continue; }
        if (c.p < c.end && *c.p == ']') return true;
        err = "expected ',' or ']'";
        return false;
    }
}

bool read_file(const char* path, std::string& text)
{
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    char buf[1 << 16];
    size_t n;
    text.clear();
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
    fclose(f);
    return true;
}

thread_local std::string g_ingest_error;

}  // namespace

extern "C" {

const char* tapir_ingest_last_error(void) { return g_ingest_error.c_str(); }

int tapir_ingest_rates(const char* const* paths, int64_t n_files, int n_threads, double correction, int64_t* offsets,
                       double* rates, int64_t capacity, int64_t* needed)
{
    if (n_files < 0 || (n_files > 0 && !paths) || !offsets || !needed) {
        g_ingest_error = "bad argument";
        return TAPIR_E_INVALID;
    }
    if (n_threads < 1) n_threads = 1;
    if (int64_t(n_threads) > n_files) n_threads = int(std::max<int64_t>(1, n_files));
    std::vector<std::vector<double>> per_file(static_cast<size_t>(n_files));
    std::atomic<int64_t> next{0};
    std::atomic<int64_t> bad{-1};
    std::vector<std::string> errors(static_cast<size_t>(n_threads));
    auto work = [&](int tid) {
        std::string text, err;
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= n_files || bad.load() >= 0) return;
            if (!read_file(paths[i], text)) {
                errors[tid] = std::string("Cannot open ") + paths[i];
                bad.store(i);
                return;
            }
            if (!parse_rates(text, per_file[size_t(i)], err)) {
                errors[tid] = std::string(paths[i]) + ": " + err;
                bad.store(i);
                return;
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    if (bad.load() >= 0) {
        for (const std::string& e : errors)
            if (!e.empty()) { g_ingest_error = e; break; }
        return TAPIR_E_INVALID;
    }
    int64_t total = 0;
    offsets[0] = 0;
    for (int64_t i = 0; i < n_files; ++i) {
        total += int64_t(per_file[size_t(i)].size());
        offsets[i + 1] = total;
    }
    *needed = total;
    if (total > capacity || (total > 0 && !rates)) return TAPIR_OK;  // caller re-calls with a buffer of *needed doubles
    for (int64_t i = 0; i < n_files; ++i) {
        const std::vector<double>& v = per_file[size_t(i)];
        double* dst = rates + offsets[i];
        for (size_t k = 0; k < v.size(); ++k) dst[k] = v[k] / correction;
    }
    return TAPIR_OK;
}

}  // extern "C"
