// csv_rules.h -- the serial loader's field rules as host+device inline functions, shared by the
// host ingest (ingest.cpp) and the GPU ingest (ingest_gpu.cu) so both follow ONE statement of
//   parseCSVField      engine/serial/buildEngine-serial.c:111-151
//   getRecordFromLine  engine/serial/buildEngine-serial.c:159-221 (strtoull / atoi / strcasecmp rules)
#pragma once

#include <cstdint>

#if defined(__CUDACC__)
#define QPE_HD __host__ __device__ __forceinline__
#else
#define QPE_HD inline
#endif

namespace qpe {
namespace csv {

// schema facts needed on the device (same order as kCols / `record`): type codes 0=u64 1=int 2=text 3=bool
QPE_HD int col_type(int c) {
    const int t[12] = {0, 2, 2, 2, 1, 2, 3, 2, 1, 2, 2, 1};
    return t[c];
}
QPE_HD unsigned col_field_bytes(int c) {  // sizeof the member of `record` (logType.h:11-24)
    const unsigned b[12] = {8, 512, 100, 20, 4, 30, 1, 200, 4, 50, 100, 4};
    return b[c];
}

QPE_HD bool is_end(char c) { return c == '\0' || c == '\n' || c == '\r'; }

// Parse one field starting at cur (bounded by end).  The unescaped text is written to out[0..cap)
// (out may be null: measure only); *len receives the full unescaped length.  Returns false when the
// field is ABSENT (it starts at \0 \n \r or at the end of the chunk): the record member stays zero.
// Rules: a leading '"' opens quoted mode; inside it '""' is a literal quote and a single '"' only
// leaves quoted mode (what follows is kept); outside it ',' ends the field; \0 \n \r end it anywhere.
QPE_HD bool next_field(const char *&cur, const char *end, char *out, int cap, int *len) {
    const char *s = cur;
    if (s >= end || is_end(*s)) return false;
    int n = 0;
    bool quoted = false;
    if (*s == '"') {
        quoted = true;
        ++s;
    }
    while (s < end && !is_end(*s)) {
        char c;
        if (quoted) {
            if (*s == '"') {
                if (s + 1 < end && s[1] == '"') {
                    c = '"';
                    s += 2;
                } else {
                    quoted = false;
                    ++s;
                    continue;
                }
            } else {
                c = *s++;
            }
        } else {
            if (*s == ',') {
                ++s;
                break;
            }
            c = *s++;
        }
        if (out && n < cap) out[n] = c;
        ++n;
    }
    *len = n;
    cur = s;
    return true;
}

QPE_HD bool is_space(char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

// strtoull(s, NULL, 10) on the first n chars: leading white space, optional sign, digits;
// overflow clamps to ULLONG_MAX; a '-' negates (modulo 2^64) unless it overflowed.
QPE_HD unsigned long long parse_u64(const char *s, int n) {
    int i = 0;
    while (i < n && is_space(s[i])) ++i;
    bool neg = false;
    if (i < n && (s[i] == '+' || s[i] == '-')) neg = s[i++] == '-';
    unsigned long long v = 0;
    bool over = false;
    for (; i < n && s[i] >= '0' && s[i] <= '9'; ++i) {
        const unsigned d = static_cast<unsigned>(s[i] - '0');
        if (v > (0xffffffffffffffffull - d) / 10ull) over = true;
        v = v * 10ull + d;
    }
    if (over) return 0xffffffffffffffffull;
    return neg ? (0ull - v) : v;
}

// atoi(s) == (int) strtol(s, NULL, 10): white space, sign, digits, clamp to LONG_MIN / LONG_MAX, truncate
QPE_HD int parse_i32(const char *s, int n) {
    int i = 0;
    while (i < n && is_space(s[i])) ++i;
    bool neg = false;
    if (i < n && (s[i] == '+' || s[i] == '-')) neg = s[i++] == '-';
    unsigned long long v = 0;
    bool over = false;
    const unsigned long long lim = neg ? 0x8000000000000000ull : 0x7fffffffffffffffull;
    for (; i < n && s[i] >= '0' && s[i] <= '9'; ++i) {
        const unsigned d = static_cast<unsigned>(s[i] - '0');
        if (over || v > (lim - d) / 10ull) {
            over = true;
            continue;
        }
        v = v * 10ull + d;
    }
    long long r;
    if (over)
        r = neg ? static_cast<long long>(0x8000000000000000ull) : 0x7fffffffffffffffll;
    else
        r = neg ? static_cast<long long>(0ull - v) : static_cast<long long>(v);
    return static_cast<int>(r);
}

// strcasecmp(tok, "true") == 0 || strcmp(tok, "1") == 0
QPE_HD bool parse_bool(const char *s, int n) {
    if (n == 1) return s[0] == '1';
    if (n != 4) return false;
    return (s[0] == 't' || s[0] == 'T') && (s[1] == 'r' || s[1] == 'R') && (s[2] == 'u' || s[2] == 'U') &&
           (s[3] == 'e' || s[3] == 'E');
}

}  // namespace csv
}  // namespace qpe
