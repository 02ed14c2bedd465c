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
#include <new>
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
            if (*p == '\\') {
                if (p + 1 >= end) return fail("unterminated string");
                ++p;
            }
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
    // after a member's value: consume the remaining members of the enclosing object up to and including its '}'
    bool finish_object()
    {
        for (;;) {
            ws();
            if (p < end && *p == '}') { ++p; return true; }
            if (p >= end || *p != ',') return fail("expected ',' or '}'");
            ++p;
            ws();
            if (!str(nullptr)) return false;
            ws();
            if (p >= end || *p != ':') return fail("expected ':'");
            ++p;
            if (!skip()) return false;
        }
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
    // the rest of the document is walked to its end as json.load does: a truncated or corrupt file is an error
    auto rest = [&]() {
        ++c.p;                                             // the ']' that closes the list
        if (!c.finish_object() || !c.finish_object()) { err = c.err; return false; }   // "sites", then the document
        c.ws();
        if (c.p != c.end) { err = "trailing data after the JSON document"; return false; }
        return true;
    };
    if (c.p < c.end && *c.p == ']') return rest();
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
        if (c.p < c.end && *c.p == ']') return rest();
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

}  // extern "C" (reopened below)

struct tapir_ingest {
    std::vector<std::vector<double>> per_file;
};

extern "C" {

int tapir_ingest_open(const char* const* paths, int64_t n_files, int n_threads, tapir_ingest** handle, int64_t* offsets,
                      int64_t* needed)
{
    if (handle) *handle = nullptr;
    if (n_files < 0 || (n_files > 0 && !paths) || !offsets || !needed || !handle) {
        g_ingest_error = "bad argument";
        return TAPIR_E_INVALID;
    }
    if (n_threads < 1) n_threads = 1;
    if (int64_t(n_threads) > n_files) n_threads = int(std::max<int64_t>(1, n_files));
    tapir_ingest* h = new (std::nothrow) tapir_ingest();
    if (!h) {
        g_ingest_error = "out of host memory";
        return TAPIR_E_NOMEM;
    }
    std::vector<std::vector<double>>& per_file = h->per_file;
    per_file.resize(static_cast<size_t>(n_files));
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
        delete h;
        return TAPIR_E_INVALID;
    }
    int64_t total = 0;
    offsets[0] = 0;
    for (int64_t i = 0; i < n_files; ++i) {
        total += int64_t(per_file[size_t(i)].size());
        offsets[i + 1] = total;
    }
    *needed = total;
    *handle = h;
    return TAPIR_OK;
}

int tapir_ingest_copy(const tapir_ingest* h, double correction, double* rates, int64_t capacity)
{
    if (!h) {
        g_ingest_error = "null handle";
        return TAPIR_E_INVALID;
    }
    int64_t total = 0;
    for (const auto& v : h->per_file) total += int64_t(v.size());
    if (total > capacity || (total > 0 && !rates)) {
        g_ingest_error = "rates buffer too small";
        return TAPIR_E_INVALID;
    }
    double* dst = rates;
    for (const auto& v : h->per_file) {
        for (size_t k = 0; k < v.size(); ++k) dst[k] = v[k] / correction;
        dst += v.size();
    }
    return TAPIR_OK;
}

int tapir_ingest_close(tapir_ingest* h)
{
    delete h;
    return TAPIR_OK;
}

int tapir_ingest_rates(const char* const* paths, int64_t n_files, int n_threads, double correction, int64_t* offsets,
                       double* rates, int64_t capacity, int64_t* needed)
{
    tapir_ingest* h = nullptr;
    int rc = tapir_ingest_open(paths, n_files, n_threads, &h, offsets, needed);
    if (rc != TAPIR_OK) return rc;
    if (*needed <= capacity && (*needed == 0 || rates)) rc = tapir_ingest_copy(h, correction, rates, capacity);
    tapir_ingest_close(h);
    return rc;   // *needed > capacity: the caller re-calls with a buffer of *needed doubles (or uses open / copy / close)
}

}  // extern "C"
