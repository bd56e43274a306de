// ingest.cpp -- CSV -> host columnar staging, with the serial loader's exact field rules.
//
// Restates engine/serial/buildEngine-serial.c:
//   getAllRecordsFromFile :70-108   one row per fgets(line, 1024) chunk after the first (header)
//                                   chunk -- so a physical line of >= 1023 characters becomes
//                                   several rows and a blank line becomes an all-zero row;
//   parseCSVField         :111-151  a field ends at ',' (outside quotes) or at \0 \n \r; a leading
//                                   '"' opens a quoted field, '""' inside is a literal quote, the
//                                   closing quote only leaves quoted mode (text after it is kept);
//                                   a field that STARTS at \0 \n \r is absent (member stays zero);
//   getRecordFromLine     :159-221  strtoull / atoi on numeric fields, strncpy into the fixed
//                                   char arrays, sudo_used = strcasecmp(tok,"true")==0 || tok=="1".
// One deliberate difference, outside the reference's defined behaviour: a string field as long
// as (or longer than) its array is stored truncated to size-1 characters WITH a terminator; the
// reference stores it unterminated and every later strcmp on it reads into the next member.

#include <cstdlib>
#include <cstring>
#include <strings.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <fcntl.h>
#include <unistd.h>

#include "qpe_internal.h"
#include "csv_rules.h"
#include "buildEngine-gpu.h"

namespace qpe {

namespace {

// one field through the shared rules (csv_rules.h); text is NUL-terminated in buf
inline bool next_field(const char *&cur, const char *end, char *buf, size_t *len_out) {
    int len = 0;
    if (!csv::next_field(cur, end, buf, 1031, &len)) return false;
    if (len > 1031) len = 1031;
    buf[len] = '\0';
    *len_out = static_cast<size_t>(len);
    return true;
}

inline void put_str(char *dst, size_t cap, const char *src, size_t len) {
    if (len > cap - 1) len = cap - 1;
    std::memcpy(dst, src, len);  // dst is pre-zeroed: the rest is the NUL padding strncpy would write
}

}  // namespace

// chunk = what one fgets(line, 1024) call would have returned (without requiring a terminator)
void parse_csv_chunk(const char *chunk, size_t len, record *r) {
    std::memset(r, 0, sizeof(record));  // calloc (:160)
    char buf[1032];
    const char *cur = chunk;
    const char *end = chunk + len;
    size_t n = 0;
    if (next_field(cur, end, buf, &n)) r->command_id = std::strtoull(buf, nullptr, 10);
    if (next_field(cur, end, buf, &n)) put_str(r->raw_command, sizeof r->raw_command, buf, n);
    if (next_field(cur, end, buf, &n)) put_str(r->base_command, sizeof r->base_command, buf, n);
    if (next_field(cur, end, buf, &n)) put_str(r->shell_type, sizeof r->shell_type, buf, n);
    if (next_field(cur, end, buf, &n)) r->exit_code = std::atoi(buf);
    if (next_field(cur, end, buf, &n)) put_str(r->timestamp, sizeof r->timestamp, buf, n);
    if (next_field(cur, end, buf, &n)) r->sudo_used = (strcasecmp(buf, "true") == 0 || std::strcmp(buf, "1") == 0);
    if (next_field(cur, end, buf, &n)) put_str(r->working_directory, sizeof r->working_directory, buf, n);
    if (next_field(cur, end, buf, &n)) r->user_id = std::atoi(buf);
    if (next_field(cur, end, buf, &n)) put_str(r->user_name, sizeof r->user_name, buf, n);
    if (next_field(cur, end, buf, &n)) put_str(r->host_name, sizeof r->host_name, buf, n);
    if (next_field(cur, end, buf, &n)) r->risk_level = std::atoi(buf);
}

// length of the next fgets(…, 1024) chunk starting at p
static inline size_t fgets_chunk(const char *p, const char *end) {
    const size_t room = static_cast<size_t>(end - p);
    const size_t lim = room < 1023 ? room : 1023;
    const void *nl = std::memchr(p, '\n', lim);
    return nl ? static_cast<size_t>(static_cast<const char *>(nl) - p) + 1 : lim;
}

struct MappedFile {
    const char *p = nullptr;
    size_t n = 0;
    int fd = -1;
    bool open(const char *path) {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0) return false;
        n = static_cast<size_t>(st.st_size);
        if (n == 0) {
            p = "";
            return true;
        }
        void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) return false;
        p = static_cast<const char *>(m);
        return true;
    }
    ~MappedFile() {
        if (p && n) munmap(const_cast<char *>(p), n);
        if (fd >= 0) ::close(fd);
    }
};

// returns -1 when the file cannot be opened (the reference prints and returns NULL, :72-76)
int64_t load_csv_columns(const char *path, HostColumns *out) {
    MappedFile f;
    if (!f.open(path)) {
        std::fprintf(stderr, "Error opening file: %s\n", path);
        return -1;
    }
    out->init_widths_minimal();
    const char *p = f.p, *end = f.p + f.n;
    // rough row estimate for reservation
    out->reserve_rows(static_cast<int64_t>(f.n / 96) + 16);
    bool first = true;
    record r;
    while (p < end) {
        const size_t len = fgets_chunk(p, end);
        if (first) {
            first = false;  // header chunk skipped (:86-89)
        } else {
            parse_csv_chunk(p, len, &r);
            out->append_record(r);
        }
        p += len;
    }
    return out->n;
}

// ------------------------------------------------------------------------------------------
// HostColumns
// ------------------------------------------------------------------------------------------
void HostColumns::init_widths_minimal() {
    n = 0;
    for (int c = 0; c < NUM_COLS; ++c) {
        switch (kCols[c].type) {
            case T_U64: width[c] = 8; break;
            case T_I32: width[c] = 4; break;
            case T_BOOL: width[c] = 1; break;
            default: width[c] = 16; break;
        }
        data[c].clear();
    }
}

void HostColumns::reserve_rows(int64_t rows) {
    for (int c = 0; c < NUM_COLS; ++c) data[c].reserve(static_cast<size_t>(rows) * width[c]);
}

void HostColumns::widen(int c, uint32_t new_width) {
    if (new_width <= width[c]) return;
    std::vector<uint8_t> nd(static_cast<size_t>(n) * new_width, 0);
    for (int64_t i = 0; i < n; ++i)
        std::memcpy(nd.data() + static_cast<size_t>(i) * new_width, data[c].data() + static_cast<size_t>(i) * width[c],
                    width[c]);
    const size_t cap = data[c].capacity() / width[c] * new_width;
    data[c].swap(nd);
    data[c].reserve(cap);
    width[c] = new_width;
}

void HostColumns::append_record(const record &r) {
    const uint8_t *rb = reinterpret_cast<const uint8_t *>(&r);
    for (int c = 0; c < NUM_COLS; ++c) {
        const uint8_t *src = rb + kCols[c].rec_offset;
        if (kCols[c].type == T_STR) {
            const size_t len = strnlen(reinterpret_cast<const char *>(src), kCols[c].field_bytes - 1);
            const uint32_t need = round_up16(static_cast<uint32_t>(len) + 1);
            if (need > width[c]) widen(c, need);
            const size_t old = data[c].size();
            data[c].resize(old + width[c], 0);
            std::memcpy(data[c].data() + old, src, len);
        } else if (kCols[c].type == T_BOOL) {
            data[c].push_back(r.sudo_used ? 1 : 0);
        } else {
            const size_t old = data[c].size();
            data[c].resize(old + width[c]);
            std::memcpy(data[c].data() + old, src, width[c]);
        }
    }
    ++n;
}

}  // namespace qpe

// ------------------------------------------------------------------------------------------
// C-ABI build-side entry points (include/buildEngine-gpu.h)
// ------------------------------------------------------------------------------------------
extern "C" {

record *getRecordFromLineGPU(char *line) {
    record *r = static_cast<record *>(std::malloc(sizeof(record)));
    if (!r) {
        std::fprintf(stderr, "Memory allocation failed\n");
        return nullptr;
    }
    qpe::parse_csv_chunk(line, std::strlen(line), r);
    return r;
}

record **getAllRecordsFromFileGPU(const char *filepath, int *num_records) {
    qpe::MappedFile f;
    if (!f.open(filepath)) {
        std::fprintf(stderr, "Error opening file: %s\n", filepath);
        return nullptr;
    }
    record **rows = nullptr;
    size_t cap = 0, cnt = 0;
    const char *p = f.p, *end = f.p + f.n;
    bool first = true;
    while (p < end) {
        const size_t len = qpe::fgets_chunk(p, end);
        if (first) {
            first = false;
        } else {
            if (cnt == cap) {
                cap = cap ? cap * 2 : 1024;
                rows = static_cast<record **>(std::realloc(rows, cap * sizeof(record *)));
                if (!rows) {
                    std::fprintf(stderr, "Memory allocation failed\n");
                    return nullptr;
                }
            }
            rows[cnt] = static_cast<record *>(std::malloc(sizeof(record)));
            qpe::parse_csv_chunk(p, len, rows[cnt]);
            ++cnt;
        }
        p += len;
    }
    if (num_records) *num_records = static_cast<int>(cnt);
    return rows;
}

}  // extern "C"
