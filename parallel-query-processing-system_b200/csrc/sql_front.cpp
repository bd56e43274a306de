// sql_front.cpp -- SQL text -> whereClauseS list -> engine call -> printed result.
//
// A restatement (not a copy) of the reference's front end so that libqpegpu.so can be driven
// with SQL text where the reference tree is absent (tests on the GPU box, bench.py, QPEGPU):
//   tokenizer   tokenizer/src/tokenizer.c:8-113     -> Lexer
//   parser      tokenizer/src/tokenizer.c:116-300   -> Parser  (same grammar, same quirks)
//   bridge      connectEngine.c:65-113, :125-245    -> build_where / qpe_sql_run
//   printer     engine/printHelper.c:35-130         -> print_result
// Inside the reference tree the maintainer links the reference's own tokenizer.c /
// connectEngine.c / printHelper.c against the *GPU entry points instead;
// oracle/Makefile's `refdriver` target does exactly that and tests diff the two front ends.
//
// Quirks kept on purpose: AND is matched by token TEXT only (must be upper
// case), OR / TRUE / FALSE are keywords (any case, stored upper-cased); no operator precedence;
// '#' and other unknown characters are skipped; "--" starts a comment; a 5th condition at one
// nesting level overwrites the condition counter (logic_ops[4] aliases num_conditions in the
// reference's struct, include/sql.h:66-67); numbers are digit runs (a leading '-' is dropped).
// Where the reference reads uninitialised memory (tokens past EOF) this front end reads EOF.

#include <cctype>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <memory>
#include <string>
#include <strings.h>
#include <vector>

#include "qpe_internal.h"
#include "executeEngine-gpu.h"
#include "qpe_gpu.h"

namespace {

constexpr int kMaxTokens = 100;      // connectEngine.h:13
constexpr size_t kTokenChars = 255;  // sql.h:44  (char value[256])

enum TokKind { TK_KEYWORD, TK_IDENT, TK_SYMBOL, TK_STRING, TK_NUMBER, TK_EOF };
struct Tok {
    TokKind kind = TK_EOF;
    std::string text;
};

const char *const kKeywords[] = {"SELECT", "FROM",  "WHERE",    "ORDER",  "BY",   "DESC",   "OR",
                                 "TRUE",   "FALSE", "DESCRIBE", "INSERT", "INTO", "VALUES", "DELETE"};

std::vector<Tok> lex(const char *in) {
    std::vector<Tok> out;
    size_t p = 0;
    auto push = [&](TokKind k, std::string s) {
        if (s.size() > kTokenChars) s.resize(kTokenChars);
        Tok t;
        t.kind = k;
        t.text = std::move(s);
        out.push_back(std::move(t));
    };
    while (in[p] && static_cast<int>(out.size()) < kMaxTokens - 1) {
        while (std::isspace(static_cast<unsigned char>(in[p]))) ++p;
        if (!in[p]) break;
        const char ch = in[p];
        if (ch == '-' && in[p + 1] == '-') {  // comment to end of line
            while (in[p] && in[p] != '\n') ++p;
            continue;
        }
        if (std::strchr(";,()*=", ch)) {
            push(TK_SYMBOL, std::string(1, ch));
            ++p;
            continue;
        }
        if (ch == '>' || ch == '<' || ch == '!') {
            std::string s(1, ch);
            if (in[p + 1] == '=') s.push_back('=');
            p += s.size();
            push(TK_SYMBOL, s);
            continue;
        }
        if (ch == '"' || ch == '\'') {
            const size_t start = ++p;
            while (in[p] && in[p] != ch) ++p;
            push(TK_STRING, std::string(in + start, p - start));
            if (in[p] == ch) ++p;
            continue;
        }
        if (std::isalnum(static_cast<unsigned char>(ch)) || ch == '_') {
            const size_t start = p;
            if (std::isdigit(static_cast<unsigned char>(ch))) {
                while (std::isdigit(static_cast<unsigned char>(in[p]))) ++p;
                if (!std::isalpha(static_cast<unsigned char>(in[p]))) {
                    push(TK_NUMBER, std::string(in + start, p - start));
                    continue;
                }  // digits followed by a letter: falls through as an identifier
            }
            while (std::isalnum(static_cast<unsigned char>(in[p])) || in[p] == '_') ++p;
            std::string word(in + start, p - start);
            std::string upper = word;
            for (char &c : upper) c = static_cast<char>(std::toupper(static_cast<unsigned char>(c)));
            bool kw = false;
            for (const char *k : kKeywords)
                if (upper == k) kw = true;
            if (kw)
                push(TK_KEYWORD, upper);
            else
                push(TK_IDENT, word);
            continue;
        }
        ++p;  // anything else ('#', '.', '-', ...) is dropped
    }
    Tok eof;
    eof.kind = TK_EOF;
    out.push_back(eof);
    return out;
}

// ---------------------------------------------------------------------------------------------
// parse tree (same capacities as include/sql.h:49-74)
// ---------------------------------------------------------------------------------------------
enum Cmd { CMD_NONE_, CMD_DESCRIBE_, CMD_SELECT_, CMD_INSERT_, CMD_DELETE_, CMD_UNKNOWN_ };
enum Op { OP_NONE_, OP_EQ_, OP_NEQ_, OP_GT_, OP_LT_, OP_GTE_, OP_LTE_ };
enum Logic { LOGIC_NONE_ = 0, LOGIC_AND_ = 1, LOGIC_OR_ = 2 };

struct Parsed;
struct Cond {
    std::string column;  // <= 63 chars
    Op op = OP_NONE_;
    std::string value;   // <= 255 chars
    bool is_numeric = false;
    bool is_nested = false;
    std::shared_ptr<Parsed> nested;
};
struct Parsed {
    Cmd command = CMD_NONE_;
    std::string table;
    std::vector<std::string> columns;  // <= 10
    bool select_all = false;
    Cond conditions[5];
    Logic logic_ops[4] = {LOGIC_NONE_, LOGIC_NONE_, LOGIC_NONE_, LOGIC_NONE_};
    int num_conditions = 0;
    std::vector<std::string> insert_values;  // stored <= 15
    int num_values = 0;
};

struct Parser {
    const std::vector<Tok> &t;
    size_t i = 0;
    explicit Parser(const std::vector<Tok> &toks) : t(toks) {}
    const Tok &cur() const { return t[i < t.size() ? i : t.size() - 1]; }  // past the end == EOF
    bool is(const char *s) const { return cur().text == s; }               // the reference compares TEXT only
    void next() {
        if (i < t.size()) ++i;
    }

    void conditions(Parsed *sql) {
        while (cur().kind != TK_EOF && !is("ORDER") && !is(";") && !is(")")) {
            if (sql->num_conditions >= 5 || sql->num_conditions < 0) break;
            Cond *c = &sql->conditions[sql->num_conditions];
            c->is_nested = false;
            c->nested.reset();
            if (is("(")) {
                next();
                c->is_nested = true;
                c->nested = std::make_shared<Parsed>();
                conditions(c->nested.get());
                if (is(")")) next();
            } else {
                if (cur().kind == TK_IDENT) {
                    c->column = cur().text.substr(0, 63);
                    next();
                }
                if (is("=")) c->op = OP_EQ_;
                else if (is("!=")) c->op = OP_NEQ_;
                else if (is(">")) c->op = OP_GT_;
                else if (is("<")) c->op = OP_LT_;
                else if (is(">=")) c->op = OP_GTE_;
                else if (is("<=")) c->op = OP_LTE_;
                else c->op = OP_NONE_;
                next();  // consumed whatever it was
                if (cur().kind == TK_STRING) {
                    c->value = cur().text;
                    c->is_numeric = false;
                    next();
                } else if (cur().kind == TK_NUMBER) {
                    c->value = cur().text;
                    c->is_numeric = true;
                    next();
                } else if (cur().kind == TK_KEYWORD && (is("TRUE") || is("FALSE"))) {
                    c->value = cur().text;
                    c->is_numeric = false;
                    next();
                }
            }
            sql->num_conditions++;
            Logic lg = LOGIC_NONE_;
            if (is("AND")) {
                lg = LOGIC_AND_;
                next();
            } else if (is("OR")) {
                lg = LOGIC_OR_;
                next();
            }
            const int slot = sql->num_conditions - 1;
            if (slot < 4)
                sql->logic_ops[slot] = lg;
            else
                sql->num_conditions = static_cast<int>(lg);  // logic_ops[4] IS num_conditions in the reference
        }
    }

    Parsed parse() {
        Parsed sql;
        if (cur().kind != TK_KEYWORD) return sql;
        if (is("DESCRIBE")) {
            sql.command = CMD_DESCRIBE_;
            next();
            if (cur().kind == TK_IDENT) sql.table = cur().text;
        } else if (is("SELECT")) {
            sql.command = CMD_SELECT_;
            next();
            while (cur().kind != TK_EOF) {
                const size_t before = i;
                if (is("*")) {
                    sql.select_all = true;
                    next();
                } else if (cur().kind == TK_IDENT) {
                    if (sql.columns.size() < 10) sql.columns.push_back(cur().text.substr(0, 63));
                    next();
                }
                if (is(",")) {
                    next();
                    continue;
                }
                if (is("FROM")) break;
                if (cur().kind == TK_EOF) break;
                if (i == before) next();  // the reference would spin forever here; skip the token instead
            }
            if (is("FROM")) {
                next();
                if (cur().kind == TK_IDENT) {
                    sql.table = cur().text;
                    next();
                }
            }
            if (is("WHERE")) {
                next();
                conditions(&sql);
            }
            // ORDER BY is parsed and then ignored by every engine (tokenizer.c:244-260)
        } else if (is("INSERT")) {
            sql.command = CMD_INSERT_;
            next();
            if (is("INTO")) next();
            if (cur().kind == TK_IDENT) {
                sql.table = cur().text;
                next();
            }
            if (is("VALUES")) next();
            if (is("(")) next();
            while (cur().kind != TK_EOF && !is(")")) {
                if (is(",")) {
                    next();
                    continue;
                }
                if (sql.insert_values.size() < 15) sql.insert_values.push_back(cur().text);
                sql.num_values++;
                next();
            }
        } else if (is("DELETE")) {
            sql.command = CMD_DELETE_;
            next();
            if (is("FROM")) next();
            if (cur().kind == TK_IDENT) {
                sql.table = cur().text;
                next();
            }
            if (is("WHERE")) {
                next();
                conditions(&sql);
            }
        } else {
            sql.command = CMD_UNKNOWN_;
        }
        return sql;
    }
};

const char *op_text(Op op) {
    switch (op) {
        case OP_EQ_: return "=";
        case OP_NEQ_: return "!=";
        case OP_GT_: return ">";
        case OP_LT_: return "<";
        case OP_GTE_: return ">=";
        case OP_LTE_: return "<=";
        default: return "=";  // connectEngine.c:33
    }
}

// ParsedSQL -> whereClauseS list (connectEngine.c:65-113).  Nodes live in `pool`; strings are
// borrowed from the parse tree, which must outlive the list.
struct whereClauseS *build_where(const Parsed &p, std::vector<std::unique_ptr<struct whereClauseS>> *pool) {
    if (p.num_conditions <= 0) return nullptr;
    struct whereClauseS *head = nullptr, *tail = nullptr;
    for (int k = 0; k < p.num_conditions && k < 5; ++k) {
        pool->emplace_back(new whereClauseS());
        struct whereClauseS *n = pool->back().get();
        std::memset(n, 0, sizeof(*n));
        const Cond &c = p.conditions[k];
        if (c.is_nested && c.nested) {
            n->sub = build_where(*c.nested, pool);
        } else {
            n->attribute = c.column.c_str();
            n->op_ = op_text(c.op);
            n->value = c.value.c_str();
            n->value_type = c.is_numeric ? 0 : 1;
        }
        if (k < p.num_conditions - 1 && k < 4)
            n->logical_op = p.logic_ops[k] == LOGIC_OR_ ? "OR" : "AND";  // NONE maps to AND (:44)
        if (!head)
            head = n;
        else
            tail->next = n;
        tail = n;
    }
    return head;
}

// ---------------------------------------------------------------------------------------------
// output sink: FILE* or a growing string
// ---------------------------------------------------------------------------------------------
struct Sink {
    FILE *f = nullptr;
    std::string *s = nullptr;
    void printf(const char *fmt, ...) __attribute__((format(printf, 2, 3))) {
        va_list ap;
        va_start(ap, fmt);
        if (f) {
            std::vfprintf(f, fmt, ap);
        } else {
            va_list ap2;
            va_copy(ap2, ap);
            const int n = std::vsnprintf(nullptr, 0, fmt, ap2);
            va_end(ap2);
            if (n > 0) {
                const size_t old = s->size();
                s->resize(old + static_cast<size_t>(n) + 1);
                std::vsnprintf(&(*s)[old], static_cast<size_t>(n) + 1, fmt, ap);
                s->resize(old + static_cast<size_t>(n));
            }
        }
        va_end(ap);
    }
    void put(const std::string &str) {
        if (f)
            std::fwrite(str.data(), 1, str.size(), f);
        else
            s->append(str);
    }
};

// printTable (engine/printHelper.c:35-130): frame sized by the header and the PRINTED rows only
void print_result(Sink &out, const struct resultSetS *r, int limit) {
    if (r == nullptr || r->data == nullptr) {
        out.printf("No data found.\n");
        return;
    }
    int shown = r->numRecords;
    if (limit > 0 && limit < shown) shown = limit;
    const int nc = r->numColumns;
    std::vector<int> w(nc > 0 ? nc : 0);
    for (int j = 0; j < nc; ++j) w[j] = static_cast<int>(std::strlen(r->columnNames[j]));
    for (int i = 0; i < shown; ++i) {
        if (!r->data[i]) continue;
        for (int j = 0; j < nc; ++j) {
            if (!r->data[i][j]) continue;
            const int len = static_cast<int>(std::strlen(r->data[i][j]));
            if (len > w[j]) w[j] = len;
        }
    }
    std::string rule = "+";
    for (int j = 0; j < nc; ++j) {
        rule.append(static_cast<size_t>(w[j]) + 2, '-');
        rule.push_back('+');
    }
    rule.push_back('\n');
    out.put(rule);
    out.printf("|");
    for (int j = 0; j < nc; ++j) out.printf(" %-*s |", w[j], r->columnNames[j]);
    out.printf("\n");
    out.put(rule);
    for (int i = 0; i < shown; ++i) {
        out.printf("|");
        if (!r->data[i]) {
            out.printf(" NULL ROW |\n");
            continue;
        }
        for (int j = 0; j < nc; ++j) out.printf(" %-*s |", w[j], r->data[i][j] ? r->data[i][j] : "NULL");
        out.printf("\n");
    }
    out.put(rule);
    if (limit > 0 && r->numRecords > limit) out.printf("... (%d more records) ...\n", r->numRecords - limit);
    out.printf("Total Records: %d | Query Time: %.4f seconds\n\n", r->numRecords, r->queryTime);
}

void copy_field(char *dst, size_t cap, const std::string &src) {  // safe_copy (connectEngine.c:21-23)
    std::snprintf(dst, cap, "%.*s", static_cast<int>(cap) - 1, src.c_str());
}

double seconds_since(const std::chrono::steady_clock::time_point &t0) {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// run_test_query (connectEngine.c:125-245)
void run_statement(struct engineS *engine, const char *statement, int max_rows, Sink &out) {
    out.printf("Executing Query: %s\n", statement);
    const std::vector<Tok> toks = lex(statement);
    if (toks.size() <= 1) {
        out.printf("Tokenization failed.\n");
        return;
    }
    Parser ps(toks);
    const Parsed sql = ps.parse();
    switch (sql.command) {
        case CMD_INSERT_: {
            if (sql.num_values != 12) {
                out.printf("Error: INSERT requires exactly 12 values.\n");
                return;
            }
            const std::vector<std::string> &v = sql.insert_values;
            record r;
            std::memset(&r, 0, sizeof r);
            r.command_id = std::strtoull(v[0].c_str(), nullptr, 10);
            copy_field(r.raw_command, sizeof r.raw_command, v[1]);
            copy_field(r.base_command, sizeof r.base_command, v[2]);
            copy_field(r.shell_type, sizeof r.shell_type, v[3]);
            r.exit_code = std::atoi(v[4].c_str());
            copy_field(r.timestamp, sizeof r.timestamp, v[5]);
            r.sudo_used = (strcasecmp(v[6].c_str(), "true") == 0 || v[6] == "1");
            copy_field(r.working_directory, sizeof r.working_directory, v[7]);
            r.user_id = std::atoi(v[8].c_str());
            copy_field(r.user_name, sizeof r.user_name, v[9]);
            copy_field(r.host_name, sizeof r.host_name, v[10]);
            r.risk_level = std::atoi(v[11].c_str());
            const auto t0 = std::chrono::steady_clock::now();
            const bool ok = executeQueryInsertGPU(engine, sql.table.c_str(), &r);
            const double dt = seconds_since(t0);
            if (ok)
                out.printf("Insert successful. Execution Time: %.6f\n\n", dt);
            else
                out.printf("Insert failed. Execution Time: %.6f\n\n", dt);
            return;
        }
        case CMD_DELETE_: {
            std::vector<std::unique_ptr<struct whereClauseS>> pool;
            struct whereClauseS *wc = build_where(sql, &pool);
            const auto t0 = std::chrono::steady_clock::now();
            struct resultSetS *res = executeQueryDeleteGPU(engine, sql.table.c_str(), wc);
            const double dt = seconds_since(t0);
            if (res) {
                out.printf("Delete successful. Rows affected: %d. Execution Time: %.6f\n\n", res->numRecords, dt);
                freeResultSet(res);
            } else {
                out.printf("Delete failed. Execution Time: %.6f\n\n", dt);
            }
            return;
        }
        case CMD_SELECT_: {
            std::vector<std::unique_ptr<struct whereClauseS>> pool;
            struct whereClauseS *wc = build_where(sql, &pool);
            std::vector<const char *> items;
            if (!sql.select_all)
                for (const std::string &c : sql.columns) items.push_back(c.c_str());
            struct resultSetS *res = executeQuerySelectGPU(engine, items.empty() ? nullptr : items.data(),
                                                           static_cast<int>(items.size()), sql.table.c_str(), wc);
            print_result(out, res, max_rows);
            if (res) freeResultSet(res);
            out.printf("\n");
            return;
        }
        case CMD_NONE_:
            out.printf("No command detected.\n");
            return;
        default:
            std::fprintf(stderr, "Unsupported command.\n");
            return;
    }
}

// WHERE of a SELECT / DELETE text, kept alive together with its parse tree
struct ParsedWhere {
    std::vector<Tok> toks;
    Parsed sql;
    std::vector<std::unique_ptr<struct whereClauseS>> pool;
    struct whereClauseS *wc = nullptr;
    bool ok = false;
    explicit ParsedWhere(const char *statement) : toks(lex(statement)) {
        if (toks.size() <= 1) return;
        Parser ps(toks);
        sql = ps.parse();
        if (sql.command != CMD_SELECT_ && sql.command != CMD_DELETE_) return;
        wc = build_where(sql, &pool);
        ok = true;
    }
};

// The last statement parsed by this thread stays parsed: a driver that repeats a statement (a benchmark loop, a
// prepared query) pays the tokenizer once.  The WHERE list is only ever read by the engine.
const ParsedWhere &parse_cached(const char *statement) {
    static thread_local std::string last_text;
    static thread_local std::unique_ptr<ParsedWhere> last;
    const char *text = statement ? statement : "";
    if (!last || last_text != text) {
        last.reset(new ParsedWhere(text));
        last_text = text;
    }
    return *last;
}

}  // namespace

extern "C" {

void qpe_sql_run(struct engineS *engine, const char *statement, int max_rows, void *out_FILE) {
    Sink out;
    out.f = out_FILE ? static_cast<FILE *>(out_FILE) : stdout;
    run_statement(engine, statement, max_rows, out);
    std::fflush(out.f);
}

char *qpe_sql_run_to_text(struct engineS *engine, const char *statement, int max_rows) {
    std::string text;
    Sink out;
    out.s = &text;
    run_statement(engine, statement, max_rows, out);
    char *res = static_cast<char *>(std::malloc(text.size() + 1));
    if (!res) return nullptr;
    std::memcpy(res, text.c_str(), text.size() + 1);
    return res;
}

int qpe_sql_select_ids(struct engineS *engine, const char *statement, int flags, unsigned int **ids_out, size_t *n_out,
                       qpe_scan_stats *stats) {
    ParsedWhere pw(statement);
    if (!pw.ok) return -7;
    if (flags & QPE_SCAN_FORCE) {
        unsigned long long m = 0;
        const unsigned int *d = nullptr;
        const int rc = qpe_gpu_select_ids_device(engine, pw.wc, flags, &m, &d, stats);
        if (rc) return rc;
        unsigned int *ids = static_cast<unsigned int *>(std::malloc(sizeof(unsigned int) * (m ? m : 1)));
        if (!ids) return -3;
        if (m && d && qpe_gpu_copy_from_device(ids, d, sizeof(unsigned int) * m)) {
            std::free(ids);
            return -4;
        }
        if (ids_out) *ids_out = ids; else std::free(ids);
        if (n_out) *n_out = static_cast<size_t>(m);
        return 0;
    }
    return qpe_gpu_select_ids(engine, pw.wc, ids_out, n_out, stats);
}

int qpe_sql_select_ids_device(struct engineS *engine, const char *statement, int flags, unsigned long long *count_out,
                              const unsigned int **d_ids_out, qpe_scan_stats *stats) {
    const ParsedWhere &pw = parse_cached(statement);
    if (!pw.ok) return -7;
    return qpe_gpu_select_ids_device(engine, pw.wc, flags, count_out, d_ids_out, stats);
}

int qpe_sql_select_ids_into(struct engineS *engine, const char *statement, int flags, unsigned int *ids, size_t cap,
                            size_t *n_out, qpe_scan_stats *stats) {
    const ParsedWhere &pw = parse_cached(statement);
    if (!pw.ok) return -7;
    return qpe_gpu_select_ids_into(engine, pw.wc, flags, ids, cap, n_out, stats);
}

int qpe_sql_scan_count(struct engineS *engine, const char *statement, unsigned long long *count_out,
                       qpe_scan_stats *stats) {
    ParsedWhere pw(statement);
    if (!pw.ok) return -7;
    return qpe_gpu_scan_count(engine, pw.wc, count_out, stats);
}

int qpe_sql_select_ids_to(struct engineS *engine, const char *statement, unsigned int *dst_device,
                          unsigned long long dst_capacity, int global_ids, unsigned long long *count_out,
                          qpe_scan_stats *stats) {
    const ParsedWhere &pw = parse_cached(statement);
    if (!pw.ok) return -7;
    return qpe_gpu_select_ids_to(engine, pw.wc, dst_device, dst_capacity, global_ids, count_out, stats);
}

int qpe_sql_shard_delete(struct engineS *engine, const char *statement, unsigned long long *deleted_total_out,
                         unsigned long long *rows_total_out) {
    ParsedWhere pw(statement);
    if (!pw.ok || pw.sql.command != CMD_DELETE_) return -7;
    return qpe_shard_delete(engine, pw.wc, deleted_total_out, rows_total_out);
}

int qpe_sql_select_ids_batch(struct engineS *engine, const char *const *statements, int n_queries,
                             unsigned int **ids_out, size_t *n_out, qpe_scan_stats *stats) {
    if (n_queries < 0 || (n_queries > 0 && !statements)) return -5;
    std::vector<std::unique_ptr<ParsedWhere>> parsed;
    std::vector<struct whereClauseS *> wcs;
    for (int q = 0; q < n_queries; ++q) {
        parsed.emplace_back(new ParsedWhere(statements[q]));
        if (!parsed.back()->ok) return -7;
        wcs.push_back(parsed.back()->wc);
    }
    return qpe_gpu_select_ids_batch(engine, wcs.data(), n_queries, ids_out, n_out, stats);
}

extern "C" void qpe_gpu_trace_put(struct engineS *engine, int slot, double ms);  // capi.cu (diagnostics)

int qpe_sql_shard_select(struct engineS *engine, const char *statement, int to_host,
                         unsigned long long *counts_out, qpe_scan_stats *stats) {
    const auto t0 = std::chrono::steady_clock::now();
    const ParsedWhere &pw = parse_cached(statement);
    if (!pw.ok) return -7;
    const int rc = qpe_shard_select(engine, pw.wc, to_host, counts_out, stats);
    qpe_gpu_trace_put(engine, 5, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    return rc;
}

int qpe_sql_shard_submit(struct engineS *engine, const char *statement, int to_host) {
    const ParsedWhere &pw = parse_cached(statement);
    if (!pw.ok) return -7;
    return qpe_shard_submit(engine, pw.wc, to_host);
}

int qpe_sql_select_segments(struct engineS *engine, const char *statement, int global_ids, int *used_index_out,
                            int *n_segments_out, size_t seg_counts_out[32], long long **keys_out,
                            unsigned int **ids_out) {
    ParsedWhere pw(statement);
    if (!pw.ok) return -7;
    return qpe_gpu_select_segments(engine, pw.wc, global_ids, used_index_out, n_segments_out, seg_counts_out, keys_out,
                                   ids_out);
}

int qpe_sql_match_mask(struct engineS *engine, const char *statement, unsigned int *bitmap, size_t n_words,
                       unsigned long long *count_out, qpe_scan_stats *stats) {
    ParsedWhere pw(statement);
    if (!pw.ok) return -7;
    return qpe_gpu_match_mask(engine, pw.wc, bitmap, n_words, count_out, stats);
}

// full SELECT through the drop-in entry point; the caller releases the result with freeResultSet
struct resultSetS *qpe_sql_select(struct engineS *engine, const char *statement) {
    ParsedWhere pw(statement);
    if (!pw.ok || pw.sql.command != CMD_SELECT_) return nullptr;
    std::vector<const char *> items;
    if (!pw.sql.select_all)
        for (const std::string &c : pw.sql.columns) items.push_back(c.c_str());
    return executeQuerySelectGPU(engine, items.empty() ? nullptr : items.data(), static_cast<int>(items.size()),
                                 pw.sql.table.c_str(), pw.wc);
}

// parity aid (no device needed): the predicate program compile_where builds for a statement's WHERE, as raw bytes
// (struct Program of csrc/qpe_internal.h), for the column widths given.  Returns the number of bytes written,
// -7 if the statement does not parse, -2 if it does not compile, -5 if `cap` is too small.
long long qpe_sql_compile_program(const char *statement, const unsigned int widths[12], void *out, size_t cap) {
    ParsedWhere pw(statement ? statement : "");
    if (!pw.ok) return -7;
    if (cap < sizeof(qpe::Program)) return -5;
    uint32_t w[qpe::NUM_COLS];
    for (int c = 0; c < qpe::NUM_COLS; ++c) w[c] = widths[c];
    qpe::Program prog;
    const std::string err = qpe::compile_where(pw.wc, w, &prog, false);
    if (!err.empty()) return -2;
    std::memcpy(out, &prog, sizeof(prog));
    return static_cast<long long>(sizeof(prog));
}

// debugging / parity aid: the whereClauseS list of a statement rendered as text, e.g.
//   "sudo_used = TRUE OR ( risk_level = 5 AND shell_type = bash )"
char *qpe_sql_where_to_text(const char *statement) {
    ParsedWhere pw(statement);
    std::string s;
    struct Rec {
        static void go(const struct whereClauseS *w, std::string *s) {
            for (; w; w = w->next) {
                if (w->sub) {
                    s->append("( ");
                    go(w->sub, s);
                    s->append(" )");
                } else {
                    s->append(w->attribute ? w->attribute : "<null>");
                    s->push_back(' ');
                    s->append(w->op_ ? w->op_ : "<null>");
                    s->push_back(' ');
                    s->append(w->value ? w->value : "<null>");
                }
                if (w->next) {
                    s->push_back(' ');
                    s->append(w->logical_op ? w->logical_op : "<none>");
                    s->push_back(' ');
                }
            }
        }
    };
    if (pw.ok) Rec::go(pw.wc, &s);
    char *res = static_cast<char *>(std::malloc(s.size() + 1));
    if (res) std::memcpy(res, s.c_str(), s.size() + 1);
    return res;
}

}  // extern "C"
