// synth.cu -- synthetic command-log table generated ON THE DEVICE, straight into columns.
//
// The reference's data-generation/generate_commands.py makes ~12 k rows/s (100 M rows = 2.3 h,
// 1 B rows = 23 h), so the 100 M / 1 B-row configurations are generated here with a
// counter-based generator: every value is a pure function of (seed, global row id, draw number),
// so any shard regenerates exactly its slice of the virtual table.  The DISTRIBUTIONS are those
// of the reference script (generate_commands.py:16-41, :589-624, :627-656,
// :687-750): user population 2*sqrt(N) capped at 2000 with log-normal activity, per-user shell
// and threat level, risk level ~ exp(-0.9 (risk-1)) skewed by threat, per-template sudo
// probability, exit codes by risk, uniform host / working directory / timestamp over one year,
// 8 % chained follow-up commands.  The command TEMPLATES are a compact set of our own (the
// reference's ~170 templates are data, not behaviour); column widths come out at the canonical
// figures of SURVEY 8(d).  command_id == global row id, as in the script (:764).

#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "engine.cuh"
#include "qpe_gpu.h"

namespace qpe {

namespace {

constexpr int kMaxUsers = 2000;
constexpr int kNumTemplates = 40;
constexpr int kThreats = 5;

struct Template {
    char base[16];
    char fmt[56];   // '%' is replaced by the argument text
    uint8_t risk;   // 1..5
    uint8_t arg;    // argument kind
    uint8_t pad[2];
    float sudo_p;
};

enum ArgKind : uint8_t { A_NONE, A_PY, A_TXT, A_LOG, A_PKG, A_BRANCH, A_HOME, A_PORT, A_REMOTE, A_IMAGE, A_PATTERN };

const Template kTemplates[kNumTemplates] = {
    // risk 1
    {"ls", "ls -la", 1, A_NONE, {0, 0}, 0.00f},
    {"ls", "ls %", 1, A_HOME, {0, 0}, 0.00f},
    {"cd", "cd %/projects", 1, A_HOME, {0, 0}, 0.00f},
    {"pwd", "pwd", 1, A_NONE, {0, 0}, 0.00f},
    {"cat", "cat %", 1, A_TXT, {0, 0}, 0.01f},
    {"echo", "echo \"Hello, world\"", 1, A_NONE, {0, 0}, 0.00f},
    {"grep", "grep -rn % .", 1, A_PATTERN, {0, 0}, 0.01f},
    {"python3", "python3 %", 1, A_PY, {0, 0}, 0.00f},
    {"git", "git status", 1, A_NONE, {0, 0}, 0.00f},
    {"git", "git checkout %", 1, A_BRANCH, {0, 0}, 0.00f},
    {"vim", "vim %", 1, A_PY, {0, 0}, 0.02f},
    {"tail", "tail -n 50 %", 1, A_LOG, {0, 0}, 0.03f},
    // risk 2
    {"pip", "pip install %", 2, A_PKG, {0, 0}, 0.10f},
    {"git", "git push origin %", 2, A_BRANCH, {0, 0}, 0.00f},
    {"make", "make -j8 all", 2, A_NONE, {0, 0}, 0.02f},
    {"ssh", "ssh %", 2, A_REMOTE, {0, 0}, 0.00f},
    {"curl", "curl -s http://localhost:%/health", 2, A_PORT, {0, 0}, 0.00f},
    {"tar", "tar -czf backup.tar.gz %", 2, A_HOME, {0, 0}, 0.05f},
    {"docker", "docker run -it %", 2, A_IMAGE, {0, 0}, 0.30f},
    {"cp", "cp % /tmp/", 2, A_TXT, {0, 0}, 0.03f},
    {"mv", "mv % old.txt", 2, A_TXT, {0, 0}, 0.03f},
    {"node", "node app.js --port %", 2, A_PORT, {0, 0}, 0.00f},
    // risk 3
    {"apt", "apt install %", 3, A_PKG, {0, 0}, 0.85f},
    {"chmod", "chmod 755 %", 3, A_PY, {0, 0}, 0.20f},
    {"kill", "kill -9 %", 3, A_PORT, {0, 0}, 0.25f},
    {"scp", "scp % backup@%:/srv", 3, A_REMOTE, {0, 0}, 0.02f},
    {"systemctl", "systemctl restart nginx", 3, A_NONE, {0, 0}, 0.90f},
    {"crontab", "crontab -e", 3, A_NONE, {0, 0}, 0.15f},
    {"wget", "wget http://%/install.sh", 3, A_REMOTE, {0, 0}, 0.05f},
    {"rm", "rm %", 3, A_LOG, {0, 0}, 0.10f},
    // risk 4
    {"chmod", "chmod -R 777 %", 4, A_HOME, {0, 0}, 0.55f},
    {"chown", "chown -R root:root %", 4, A_HOME, {0, 0}, 0.95f},
    {"iptables", "iptables -F", 4, A_NONE, {0, 0}, 0.97f},
    {"nc", "nc -lvp %", 4, A_PORT, {0, 0}, 0.20f},
    {"curl", "curl http://%/x.sh | sh", 4, A_REMOTE, {0, 0}, 0.35f},
    {"passwd", "passwd root", 4, A_NONE, {0, 0}, 0.99f},
    // risk 5
    {"rm", "rm -rf %", 5, A_HOME, {0, 0}, 0.60f},
    {"dd", "dd if=/dev/zero of=/dev/sda bs=1M", 5, A_NONE, {0, 0}, 0.98f},
    {"mkfs", "mkfs.ext4 /dev/sdb1", 5, A_NONE, {0, 0}, 0.98f},
    {"rm", "rm -rf / --no-preserve-root", 5, A_NONE, {0, 0}, 0.90f},
};

struct SynthTables {
    uint32_t user_cdf[kMaxUsers];                 // cumulative activity weight, scaled to 2^32
    uint8_t user_shell[kMaxUsers];                // 0 bash, 1 zsh, 2 fish, 3 sh
    uint8_t user_threat[kMaxUsers];               // 0..4
    uint32_t tmpl_cdf[kThreats][kNumTemplates];   // per threat level
    Template tmpl[kNumTemplates];
    int n_users;
};

__device__ __host__ inline uint64_t mix64(uint64_t x) {
    x ^= x >> 30;
    x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27;
    x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}
// draw k of row `row` under `seed`
__device__ __host__ inline uint64_t draw(uint64_t seed, uint64_t row, uint32_t k) {
    return mix64(mix64(seed + 0x9e3779b97f4a7c15ull * (row + 1)) ^ (0xd1b54a32d192ed03ull * (k + 1)));
}
__device__ __host__ inline uint32_t draw_u32(uint64_t seed, uint64_t row, uint32_t k) {
    return static_cast<uint32_t>(draw(seed, row, k) >> 32);
}
__device__ __host__ inline uint32_t draw_below(uint64_t seed, uint64_t row, uint32_t k, uint32_t n) {
    return static_cast<uint32_t>((static_cast<uint64_t>(draw_u32(seed, row, k)) * n) >> 32);
}

__device__ inline int cdf_pick(const uint32_t *cdf, int n, uint32_t u) {  // first i with u < cdf[i]
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (u < cdf[mid]) hi = mid; else lo = mid + 1;
    }
    return lo;
}

struct Text {  // small fixed buffer string builder (device)
    char *p;
    int n, cap;
    __device__ void put(const char *s) {
        while (*s && n < cap - 1) p[n++] = *s++;
    }
    __device__ void putc(char c) {
        if (n < cap - 1) p[n++] = c;
    }
    __device__ void put_uint(uint32_t v, int min_digits) {
        char tmp[12];
        int k = 0;
        do {
            tmp[k++] = static_cast<char>('0' + v % 10);
            v /= 10;
        } while (v);
        while (k < min_digits) tmp[k++] = '0';
        while (k) putc(tmp[--k]);
    }
};

__device__ inline void store_padded(uint8_t *dst, uint32_t width, const char *src, int len) {
    // width is a multiple of 16 and dst is 16-byte aligned
    for (uint32_t o = 0; o < width; o += 16) {
        uint32_t w[4] = {0, 0, 0, 0};
        char *b = reinterpret_cast<char *>(w);
        for (int i = 0; i < 16; ++i) {
            const int idx = static_cast<int>(o) + i;
            b[i] = idx < len ? src[idx] : '\0';
        }
        *reinterpret_cast<uint4 *>(dst + o) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

struct SynthParams {
    uint8_t *col[NUM_COLS];
    uint32_t width[NUM_COLS];
    uint32_t mask;
    uint64_t seed;
    uint64_t row_base;
    long long n_rows;
    const SynthTables *tab;
};

__device__ const char kShells[4][8] = {"bash", "zsh", "fish", "sh"};
__device__ const char kHosts[16][16] = {"labpc-01", "labpc-02", "labpc-03", "labpc-04", "labpc-05", "labpc-06",
                                           "labpc-07", "labpc-08", "labpc-09", "labpc-10", "vm-ubuntu-01", "vm-ubuntu-02",
                                           "cs-lab-01", "cs-lab-02", "personal-laptop", "remote-ssh-01"};
__device__ const char kSubdirs[9][20] = {"", "/projects", "/projects/cs101", "/projects/cs201", "/projects/research",
                                            "/Downloads", "/Desktop", "/.config", "/Documents"};
__device__ const char kSysdirs[3][12] = {"/tmp", "/var/log", "/etc"};
__device__ const char kPkgs[6][12] = {"numpy", "pandas", "torch", "django", "flask", "matplotlib"};
__device__ const char kBranches[4][12] = {"main", "dev", "feature-x", "bugfix-y"};
__device__ const char kPorts[4][8] = {"8000", "8080", "3000", "5432"};
__device__ const char kRemotes[3][20] = {"login.cluster.edu", "github.com", "gitlab.com"};
__device__ const char kImages[4][16] = {"ubuntu:20.04", "python:3.11", "postgres:15", "nginx:latest"};
__device__ const char kPatterns[5][8] = {"TODO", "ERROR", "WARNING", "fixme", "BUG"};
__device__ const char kFollow[4][16] = {"echo \"done\"", "pwd", "ls", "echo \"OK\""};
__device__ const int kFailCodes[5] = {1, 2, 126, 127, 130};
__device__ const float kFailProb[6] = {0.f, 0.03f, 0.06f, 0.10f, 0.16f, 0.22f};

// days since 1970-01-01 -> civil date
__device__ inline void civil_from_days(long long z, int *y, int *m, int *d) {
    z += 719468;
    const long long era = (z >= 0 ? z : z - 146096) / 146097;
    const unsigned doe = static_cast<unsigned>(z - era * 146097);
    const unsigned yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    const long long yy = static_cast<long long>(yoe) + era * 400;
    const unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    const unsigned mp = (5 * doy + 2) / 153;
    *d = static_cast<int>(doy - (153 * mp + 2) / 5 + 1);
    *m = static_cast<int>(mp < 10 ? mp + 3 : mp - 9);
    *y = static_cast<int>(yy + (*m <= 2));
}

__global__ void __launch_bounds__(256) synth_kernel(const __grid_constant__ SynthParams p) {
    const SynthTables &T = *p.tab;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < p.n_rows;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const uint64_t row = p.row_base + static_cast<uint64_t>(i);
        const uint64_t s = p.seed;
        const int u = cdf_pick(T.user_cdf, T.n_users, draw_u32(s, row, 0));
        const int threat = T.user_threat[u];
        const int ti = cdf_pick(T.tmpl_cdf[threat], kNumTemplates, draw_u32(s, row, 1));
        const Template &tm = T.tmpl[ti];
        const int risk = tm.risk;
        const bool sudo = (draw_u32(s, row, 2) >> 8) * (1.0f / 16777216.0f) < tm.sudo_p;
        const uint32_t user_id = 1000u + static_cast<uint32_t>(u);

        if (p.mask & (1u << C_COMMAND_ID)) reinterpret_cast<unsigned long long *>(p.col[C_COMMAND_ID])[i] = row;
        if (p.mask & (1u << C_USER_ID)) reinterpret_cast<int *>(p.col[C_USER_ID])[i] = static_cast<int>(user_id);
        if (p.mask & (1u << C_RISK_LEVEL)) reinterpret_cast<int *>(p.col[C_RISK_LEVEL])[i] = risk;
        if (p.mask & (1u << C_SUDO_USED)) p.col[C_SUDO_USED][i] = sudo ? 1 : 0;
        if (p.mask & (1u << C_EXIT_CODE)) {
            int code = 0;
            if ((draw_u32(s, row, 3) >> 8) * (1.0f / 16777216.0f) < kFailProb[risk])
                code = kFailCodes[draw_below(s, row, 4, 5)];
            reinterpret_cast<int *>(p.col[C_EXIT_CODE])[i] = code;
        }
        char buf[160];
        if (p.mask & (1u << C_SHELL_TYPE)) {
            Text t{buf, 0, 160};
            t.put(kShells[T.user_shell[u]]);
            store_padded(p.col[C_SHELL_TYPE] + static_cast<size_t>(i) * p.width[C_SHELL_TYPE], p.width[C_SHELL_TYPE], buf, t.n);
        }
        if (p.mask & (1u << C_HOST_NAME)) {
            Text t{buf, 0, 160};
            t.put(kHosts[draw_below(s, row, 5, 16)]);
            store_padded(p.col[C_HOST_NAME] + static_cast<size_t>(i) * p.width[C_HOST_NAME], p.width[C_HOST_NAME], buf, t.n);
        }
        if (p.mask & (1u << C_USER_NAME)) {
            Text t{buf, 0, 160};
            t.put("student");
            t.put_uint(user_id, 1);
            store_padded(p.col[C_USER_NAME] + static_cast<size_t>(i) * p.width[C_USER_NAME], p.width[C_USER_NAME], buf, t.n);
        }
        if (p.mask & (1u << C_BASE_COMMAND)) {
            Text t{buf, 0, 160};
            t.put(tm.base);
            store_padded(p.col[C_BASE_COMMAND] + static_cast<size_t>(i) * p.width[C_BASE_COMMAND], p.width[C_BASE_COMMAND], buf, t.n);
        }
        if (p.mask & (1u << C_WORKING_DIRECTORY)) {
            Text t{buf, 0, 160};
            const uint32_t k = draw_below(s, row, 6, 12);
            if (k >= 9) {
                t.put(kSysdirs[k - 9]);
            } else {
                t.put("/home/student");
                t.put_uint(user_id, 1);
                t.put(kSubdirs[k]);
            }
            store_padded(p.col[C_WORKING_DIRECTORY] + static_cast<size_t>(i) * p.width[C_WORKING_DIRECTORY],
                         p.width[C_WORKING_DIRECTORY], buf, t.n);
        }
        if (p.mask & (1u << C_TIMESTAMP)) {
            // uniform over the 365 days before 2026-10-18T00:00:00Z, millisecond resolution
            const uint64_t span_ms = 365ull * 86400ull * 1000ull;
            const uint64_t off = draw(s, row, 7) % span_ms;
            const long long day0 = 20744 - 365;  // 2026-10-18 is day 20744 since 1970-01-01
            const long long day = day0 + static_cast<long long>(off / 86400000ull);
            uint32_t ms = static_cast<uint32_t>(off % 86400000ull);
            int y, m, d;
            civil_from_days(day, &y, &m, &d);
            Text t{buf, 0, 160};
            t.put_uint(static_cast<uint32_t>(y), 4); t.putc('-');
            t.put_uint(static_cast<uint32_t>(m), 2); t.putc('-');
            t.put_uint(static_cast<uint32_t>(d), 2); t.putc('T');
            t.put_uint(ms / 3600000u, 2); t.putc(':'); ms %= 3600000u;
            t.put_uint(ms / 60000u, 2); t.putc(':'); ms %= 60000u;
            t.put_uint(ms / 1000u, 2); t.putc('.');
            t.put_uint(ms % 1000u, 3); t.putc('Z');
            store_padded(p.col[C_TIMESTAMP] + static_cast<size_t>(i) * p.width[C_TIMESTAMP], p.width[C_TIMESTAMP], buf, t.n);
        }
        if (p.mask & (1u << C_RAW_COMMAND)) {
            Text t{buf, 0, 128};
            if (sudo) t.put("sudo ");
            char arg[48];
            Text a{arg, 0, 48};
            switch (tm.arg) {
                case A_PY: a.put("main"); a.put_uint(draw_below(s, row, 8, 6), 1); a.put(".py"); break;
                case A_TXT: a.put("notes"); a.put_uint(draw_below(s, row, 8, 10), 1); a.put(".txt"); break;
                case A_LOG: a.put("app"); a.put_uint(draw_below(s, row, 8, 4), 1); a.put(".log"); break;
                case A_PKG: a.put(kPkgs[draw_below(s, row, 8, 6)]); break;
                case A_BRANCH: a.put(kBranches[draw_below(s, row, 8, 4)]); break;
                case A_HOME: a.put("/home/student"); a.put_uint(user_id, 1); break;
                case A_PORT: a.put(kPorts[draw_below(s, row, 8, 4)]); break;
                case A_REMOTE: a.put(kRemotes[draw_below(s, row, 8, 3)]); break;
                case A_IMAGE: a.put(kImages[draw_below(s, row, 8, 4)]); break;
                case A_PATTERN: a.put(kPatterns[draw_below(s, row, 8, 5)]); break;
                default: break;
            }
            arg[a.n] = '\0';
            for (const char *f = tm.fmt; *f; ++f) {
                if (*f == '%') t.put(arg); else t.putc(*f);
            }
            if (risk <= 3 && (draw_u32(s, row, 9) >> 8) * (1.0f / 16777216.0f) < 0.08f) {
                t.put((draw_u32(s, row, 10) & 1u) ? " && " : " | ");
                t.put(kFollow[draw_below(s, row, 11, 4)]);
            }
            store_padded(p.col[C_RAW_COMMAND] + static_cast<size_t>(i) * p.width[C_RAW_COMMAND], p.width[C_RAW_COMMAND], buf, t.n);
        }
    }
}

// host-side population tables (generate_users, generate_commands.py:589-624)
void build_tables(uint64_t total_rows, uint64_t seed, SynthTables *T) {
    std::memset(T, 0, sizeof(*T));
    int n_users = static_cast<int>(std::fmax(10.0, std::fmin(2000.0, 2.0 * std::sqrt(static_cast<double>(total_rows)))));
    T->n_users = n_users;
    const uint64_t us = seed ^ 0x75736572735f5f5full;  // separate stream for the population
    std::vector<double> w(n_users);
    double total = 0;
    const double threat_w[kThreats] = {1.0, 0.3, 0.08, 0.02, 0.005};
    const double shell_w[4] = {0.7, 0.2, 0.05, 0.05};
    auto unit = [&](uint64_t i, uint32_t k) { return (static_cast<double>(draw(us, i, k) >> 11) + 0.5) / 9007199254740992.0; };
    auto pick = [&](const double *ws, int n, double u) {
        double sum = 0;
        for (int i = 0; i < n; ++i) sum += ws[i];
        double acc = 0;
        for (int i = 0; i < n; ++i) {
            acc += ws[i] / sum;
            if (u < acc) return i;
        }
        return n - 1;
    };
    for (int i = 0; i < n_users; ++i) {
        T->user_shell[i] = static_cast<uint8_t>(pick(shell_w, 4, unit(i, 0)));
        T->user_threat[i] = static_cast<uint8_t>(pick(threat_w, kThreats, unit(i, 1)));
        const double z = std::sqrt(-2.0 * std::log(unit(i, 2))) * std::cos(6.283185307179586 * unit(i, 3));
        w[i] = std::exp(z) * (1.0 + 0.3 * T->user_threat[i]);  // lognormal(0,1) * threat scaling
        total += w[i];
    }
    double acc = 0;
    for (int i = 0; i < n_users; ++i) {
        acc += w[i] / total;
        const double v = acc * 4294967296.0;
        T->user_cdf[i] = v >= 4294967295.0 ? 0xffffffffu : static_cast<uint32_t>(v);
    }
    T->user_cdf[n_users - 1] = 0xffffffffu;
    // template weights: exp(-0.9 (risk-1)) shared among the templates of a risk level, scaled by
    // (1 + 0.4 * threat * (risk-1))  (choose_command_template_for_user, :687-695)
    int per_risk[6] = {0};
    for (int t = 0; t < kNumTemplates; ++t) per_risk[kTemplates[t].risk]++;
    for (int th = 0; th < kThreats; ++th) {
        double tw[kNumTemplates], sum = 0;
        for (int t = 0; t < kNumTemplates; ++t) {
            const int r = kTemplates[t].risk;
            tw[t] = std::exp(-0.9 * (r - 1)) / per_risk[r] * (1.0 + 0.4 * th * (r - 1));
            sum += tw[t];
        }
        double a = 0;
        for (int t = 0; t < kNumTemplates; ++t) {
            a += tw[t] / sum;
            const double v = a * 4294967296.0;
            T->tmpl_cdf[th][t] = v >= 4294967295.0 ? 0xffffffffu : static_cast<uint32_t>(v);
        }
        T->tmpl_cdf[th][kNumTemplates - 1] = 0xffffffffu;
    }
    std::memcpy(T->tmpl, kTemplates, sizeof(kTemplates));
}

// canonical widths of SURVEY 8(d) for this generator's data
uint32_t synth_width(int c) {
    switch (c) {
        case C_COMMAND_ID: return 8;
        case C_EXIT_CODE:
        case C_USER_ID:
        case C_RISK_LEVEL: return 4;
        case C_SUDO_USED: return 1;
        case C_TIMESTAMP: return 32;
        case C_WORKING_DIRECTORY: return 48;
        case C_RAW_COMMAND: return 128;
        default: return 16;  // shell_type, base_command, user_name, host_name
    }
}

}  // namespace

bool synth_fill(GpuEngine *g, uint64_t total_rows, uint64_t row_base, uint64_t n_rows, uint64_t seed, uint32_t mask) {
    cudaSetDevice(g->device);
    std::vector<SynthTables> host(1);
    build_tables(total_rows, seed, &host[0]);
    SynthTables *d_tab = nullptr;
    if (!cuda_ok(cudaMalloc(&d_tab, sizeof(SynthTables)), "cudaMalloc synth tables")) return false;
    bool ok = cuda_ok(cudaMemcpyAsync(d_tab, host.data(), sizeof(SynthTables), cudaMemcpyHostToDevice, g->stream),
                      "upload synth tables");
    SynthParams p{};
    int64_t want = static_cast<int64_t>(n_rows) + kRowPad - 1;
    want = want / kRowPad * kRowPad + kRowPad;
    for (int c = 0; c < NUM_COLS && ok; ++c) {
        g->table.col[c] = DevColumn();
        g->table.col[c].width = synth_width(c);  // the layout is defined even for non-resident columns
        if (!(mask & (1u << c))) continue;
        ok = column_alloc(&g->table.col[c], synth_width(c), want, g->stream);
        p.col[c] = g->table.col[c].d;
        p.width[c] = g->table.col[c].width;
    }
    if (ok) {
        p.mask = mask;
        p.seed = seed;
        p.row_base = row_base;
        p.n_rows = static_cast<long long>(n_rows);
        p.tab = d_tab;
        if (n_rows > 0) {
            synth_kernel<<<148 * 8, 256, 0, g->stream>>>(p);
            ok = cuda_ok(cudaGetLastError(), "synth kernel launch");
        }
    }
    ok = cuda_ok(cudaStreamSynchronize(g->stream), "synth sync") && ok;
    cudaFree(d_tab);
    if (!ok) return false;
    g->table.n = static_cast<int64_t>(n_rows);
    g->table.row_base = row_base;
    g->head.num_records = static_cast<int>(n_rows);
    return true;
}

}  // namespace qpe

extern "C" struct engineS *qpe_gpu_engine_synth(unsigned long long total_rows, unsigned long long row_base,
                                                unsigned long long n_rows, unsigned long long seed,
                                                unsigned int column_mask, int num_indexes,
                                                const char *indexed_attributes[], const int attribute_types[]) {
    using namespace qpe;
    if (n_rows > 0x7fffffffull) {
        set_error("a shard holds at most INT_MAX rows (row ids are 32-bit, counts are int as in the reference)");
        return nullptr;
    }
    GpuEngine *g = engine_create("commands", nullptr, num_indexes);
    if (!g) return nullptr;
    if (!synth_fill(g, total_rows, row_base, n_rows, seed, column_mask & 0xfffu)) {
        engine_destroy(g);
        return nullptr;
    }
    for (int i = 0; i < num_indexes; ++i)
        if (!engine_add_index(g, indexed_attributes[i], attribute_types ? attribute_types[i] : -1))
            std::fprintf(stderr, "Failed to create index for attribute: %s\n", indexed_attributes[i]);
    return &g->head;
}
