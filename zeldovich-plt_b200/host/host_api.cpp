// Host half of the C ABI: parameter loading, power-spectrum set-up, eigenmode file
// loading, ic_* file writing and the whole `zeldovich <param_file>` flow.  The device
// half lives in csrc/zplt_api.cu; this file only talks to it through the C ABI.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/zeldovich_b200.h"
#include "host_parameters.h"
#include "host_power.h"

// error channel shared with the device half
extern "C" void zplt_set_error_(const char *msg);

static int hfail(int code, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    zplt_set_error_(buf);
    return code;
}

static void put_str(char *dst, size_t cap, const std::string &s) {
    snprintf(dst, cap, "%s", s.c_str());
}

static void to_pod(const Parameters &P, zplt_params *o) {
    memset(o, 0, sizeof(*o));
    o->boxsize = P.boxsize, o->Pk_scale = P.Pk_scale, o->separation = P.separation, o->fundamental = P.fundamental;
    o->nyquist = P.nyquist, o->k_cutoff = P.k_cutoff, o->Pk_norm = P.Pk_norm, o->Pk_sigma = P.Pk_sigma;
    o->Pk_sigma_ratio = P.Pk_sigma_ratio, o->f_cluster = P.f_cluster, o->Pk_smooth = P.Pk_smooth;
    o->Pk_powerlaw_index = P.Pk_powerlaw_index, o->z_initial = P.z_initial, o->PLT_target_z = P.PLT_target_z;
    o->f_NL = P.f_NL, o->n_s = P.n_s, o->Omega_M = P.Omega_M;
    o->ppd = P.ppd, o->np = P.np, o->cpd = P.cpd, o->numblock = P.numblock, o->qdensity = P.qdensity, o->qascii = P.qascii;
    o->qoneslab = P.qoneslab, o->seed = P.seed, o->qPk_fix_to_mean = P.qPk_fix_to_mean, o->qonemode = P.qonemode;
    for (int i = 0; i < 3; i++) o->one_mode[i] = (size_t) i < P.one_mode.size() ? P.one_mode[i] : 0;
    o->qPLT = P.qPLT, o->qPLTrescale = P.qPLTrescale, o->AllowDirectIO = P.AllowDirectIO, o->version = P.version;
    o->CornerModes = P.CornerModes;
    put_str(o->Pk_filename, sizeof(o->Pk_filename), P.Pk_filename.string());
    put_str(o->output_dir, sizeof(o->output_dir), P.output_dir.string());
    put_str(o->density_filename, sizeof(o->density_filename), P.density_filename.string());
    put_str(o->PLT_filename, sizeof(o->PLT_filename), P.PLT_filename.string());
    put_str(o->ICFormat, sizeof(o->ICFormat), P.ICFormat);
}

extern "C" int zplt_params_load(const char *param_file, zplt_params *out) {
    if (!param_file || !out) return hfail(ZPLT_EINVAL, "null argument");
    try {
        Parameters P(param_file);
        to_pod(P, out);
    } catch (const std::exception &e) {
        return hfail(ZPLT_EINVAL, "%s", e.what());
    }
    return ZPLT_OK;
}

extern "C" int zplt_icformat_code(const char *f) {
    if (!f) return -1;
    if (!strcmp(f, "RVdoubleZel")) return ZPLT_FMT_RVDOUBLEZEL;
    if (!strcmp(f, "RVZel")) return ZPLT_FMT_RVZEL;
    if (!strcmp(f, "Zeldovich")) return ZPLT_FMT_ZELDOVICH;
    if (!strcmp(f, "ZelSimple")) return ZPLT_FMT_ZELSIMPLE;
    return -1;
}

extern "C" int zplt_config_from_params(const zplt_params *p, zplt_config *c) {
    if (!p || !c) return hfail(ZPLT_EINVAL, "null argument");
    memset(c, 0, sizeof(*c));
    int fmt = zplt_icformat_code(p->ICFormat);
    if (fmt < 0) return hfail(ZPLT_EINVAL, "Error: unknown ICFormat \"%s\". Aborting.", p->ICFormat);
    if (p->qdensity < 0 || p->qdensity > 2) return hfail(ZPLT_EINVAL, "ZD_qdensity must be 0, 1 or 2");
    c->ppd          = p->ppd;
    c->boxsize      = p->boxsize;
    c->seed         = (int64_t) p->seed;  // int -> unsigned long sign-extends (reference src/power_spectrum.cpp:14)
    c->k_cutoff     = p->k_cutoff;
    c->corner_modes = p->CornerModes;
    c->qonemode     = p->qonemode;
    for (int i = 0; i < 3; i++) c->one_mode[i] = p->one_mode[i];
    c->qPLT         = p->qPLT;
    c->qPLTrescale  = p->qPLTrescale;
    c->fixed_power  = p->qPk_fix_to_mean;
    c->PLT_target_z = p->PLT_target_z;
    c->z_initial    = p->z_initial;
    c->f_cluster    = p->f_cluster;
    c->icformat     = fmt;
    c->device       = -1;
    c->rank         = 0;
    c->nranks       = 1;
    c->f_NL         = p->f_NL;
    c->n_s          = p->n_s;
    c->Omega_M      = p->Omega_M;
    return ZPLT_OK;
}

// ---------------------------------------------------------------- power -----------
struct zplt_power {
    PowerSpectrum pk;
};

static PkParams pk_params(const zplt_params *p) {
    PkParams q;
    q.boxsize = p->boxsize, q.Pk_scale = p->Pk_scale, q.Pk_norm = p->Pk_norm, q.Pk_sigma = p->Pk_sigma;
    q.Pk_sigma_ratio = p->Pk_sigma_ratio, q.Pk_smooth = p->Pk_smooth, q.qPk_fix_to_mean = p->qPk_fix_to_mean;
    q.n_s = p->n_s;
    return q;
}

extern "C" int zplt_power_create(const zplt_params *p, zplt_power **out) {
    if (!p || !out) return hfail(ZPLT_EINVAL, "null argument");
    *out = nullptr;
    zplt_power *h = new zplt_power();
    try {
        int rc;
        if (p->Pk_filename[0])
            rc = h->pk.InitFromFile(p->Pk_filename, pk_params(p));
        else
            rc = h->pk.InitFromPowerLaw(p->Pk_powerlaw_index, pk_params(p));
        if (rc) {
            delete h;
            return hfail(ZPLT_EINVAL, "could not initialise the power spectrum from \"%s\"", p->Pk_filename);
        }
    } catch (const std::exception &e) {
        delete h;
        return hfail(ZPLT_EINVAL, "%s", e.what());
    }
    *out = h;
    return ZPLT_OK;
}
extern "C" void zplt_power_destroy(zplt_power *h) { delete h; }
extern "C" int zplt_power_info(const zplt_power *h, int32_t *n, double *norm, double *sm2) {
    if (!h) return hfail(ZPLT_EINVAL, "null argument");
    if (n) *n = h->pk.is_powerlaw ? 0 : h->pk.size();
    if (norm) *norm = h->pk.normalization;
    if (sm2) *sm2 = h->pk.Pk_smooth2;
    return ZPLT_OK;
}
extern "C" int zplt_power_arrays(const zplt_power *h, double *x, double *y, double *y2) {
    if (!h || !x || !y || !y2) return hfail(ZPLT_EINVAL, "null argument");
    const int n = h->pk.size();
    memcpy(x, h->pk.x.data(), n * sizeof(double));
    memcpy(y, h->pk.y.data(), n * sizeof(double));
    memcpy(y2, h->pk.y2.data(), n * sizeof(double));
    return ZPLT_OK;
}
extern "C" double zplt_power_eval(zplt_power *h, double k) { return h ? h->pk.power(k) : 0.0; }
extern "C" double zplt_power_sigmaR(zplt_power *h, double R) {
    try {
        return h ? h->pk.sigmaR(R) : 0.0;
    } catch (const std::exception &e) {
        hfail(ZPLT_EINVAL, "%s", e.what());
        return -1.0;
    }
}
extern "C" double zplt_power_infer_Tk(zplt_power *h, double k) { return h ? h->pk.infer_Tk(k) : 0.0; }
extern "C" double zplt_power_primordial_norm(const zplt_power *h) { return h ? h->pk.primordial_norm : 0.0; }

extern "C" int zplt_power_apply(zplt_power *h, zplt_ctx *ctx) {
    if (!h || !ctx) return hfail(ZPLT_EINVAL, "null argument");
    // the scalar behind infer_Tk; only read by the device when ZD_f_NL != 0
    if (h->pk.primordial_norm > 0. && std::isfinite(h->pk.primordial_norm)) {
        if (int rc = zplt_set_primordial(ctx, h->pk.primordial_norm)) return rc;
    }
    if (h->pk.is_powerlaw) return zplt_set_power_law(ctx, h->pk.powerlaw_index, h->pk.normalization, h->pk.Pk_smooth2);
    return zplt_set_power_spline(ctx, h->pk.size(), h->pk.x.data(), h->pk.y.data(), h->pk.y2.data(), h->pk.normalization,
                                 h->pk.Pk_smooth2);
}

// ---------------------------------------------------------------- eigenmodes ------
extern "C" int zplt_load_eigenmodes_file(zplt_ctx *ctx, const char *path) {
    if (!ctx || !path) return hfail(ZPLT_EINVAL, "null argument");
    fprintf(stderr, "Using PLT eigenmodes.\n");
    std::ifstream f(path, std::ios::in | std::ios::binary | std::ios::ate);
    if (!f) return hfail(ZPLT_EINVAL, "[Error] Could not open eigenmode file \"%s\".", path);
    const std::streamoff size = f.tellg();
    f.seekg(0, std::ios::beg);
    int32_t ppd_e = 0;
    f.read((char *) &ppd_e, sizeof(ppd_e));
    if (!f || ppd_e <= 0 || ppd_e > 4096) return hfail(ZPLT_EINVAL, "[Error] Eigenmode file \"%s\" has a bad header.", path);
    const size_t nelem  = (size_t) ppd_e * ppd_e * (ppd_e / 2 + 1) * 4;
    const size_t nbytes = nelem * sizeof(double);
    if ((size_t) size != nbytes + sizeof(ppd_e))
        return hfail(ZPLT_EINVAL, "[Error] Eigenmode file \"%s\" of size %lld did not match expected size %zu from eig_vecs_ppd %d.", path,
                     (long long) size, nbytes, ppd_e);
    std::vector<double> tab(nelem);
    f.read((char *) tab.data(), nbytes);
    if (!f) return hfail(ZPLT_EINVAL, "[Error] short read on eigenmode file \"%s\".", path);
    return zplt_set_eigenmodes(ctx, ppd_e, tab.data());
}

// ---------------------------------------------------------------- ic_* files ------
static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct WriteStats {
    int64_t files = 0, bytes = 0;
    double seconds = 0;
};

// pinned host staging (device half of the library); falls back to pageable memory if pinning fails
extern "C" void *zplt_pinned_alloc_(size_t bytes);
extern "C" void zplt_pinned_free_(void *p);

struct HostBuffer {
    unsigned char *p = nullptr;
    bool pinned      = false;
    explicit HostBuffer(size_t bytes, bool pin = true) {
        if (pin) p = (unsigned char *) zplt_pinned_alloc_(bytes);
        pinned = p != nullptr;
        if (!p) p = (unsigned char *) malloc(bytes);
    }
    ~HostBuffer() {
        if (pinned)
            zplt_pinned_free_(p);
        else
            free(p);
    }
};

// One append to one ic file.  Planes of a file are contiguous in the staging buffer (ascending z), so a chunk of planes
// turns into one job per file; the jobs of a chunk are independent files and are written by several threads at once.
struct WriteJob {
    int64_t fileno;
    const unsigned char *p;
    size_t bytes;
};

static bool run_write_jobs(const fs::path &dir, const std::vector<WriteJob> &jobs, int nthreads, std::string *err) {
    std::atomic<size_t> next{0};
    std::atomic<bool> ok{true};
    std::mutex emu;
    auto worker = [&]() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= jobs.size() || !ok.load()) return;
            const WriteJob &j = jobs[i];
            fs::path fn       = dir / ("ic_" + std::to_string(j.fileno));
            FILE *fp          = fopen(fn.c_str(), "ab");  // "ab": reference src/output.cpp:208-212
            bool good         = fp && fwrite(j.p, 1, j.bytes, fp) == j.bytes;
            if (fp) good = (fclose(fp) == 0) && good;
            if (!good) {
                std::lock_guard<std::mutex> g(emu);
                if (ok.exchange(false)) *err = "cannot append to \"" + fn.string() + "\"";
            }
        }
    };
    const int nt = std::max(1, std::min<int>(nthreads, (int) jobs.size()));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; t++) pool.emplace_back(worker);
    worker();
    for (auto &t : pool) t.join();
    return ok.load();
}

// SetupOutputDir + the ZeldovichXY write loop (reference src/output.cpp:236-251, :208-212; src/zeldovich.cpp:667-682).
// qoneslab >= 0: write only that z plane (reference src/zeldovich.cpp:669-682).
// The planes come off the device in chunks into two pinned staging buffers: while the writer threads append chunk i to its
// ic files, the device emits and copies chunk i+1.  Append order per file stays ascending z: a file's planes inside a chunk
// are one job, and a chunk is finished before the next one is handed to the writers.
// open() once, then write() for consecutive ranges of global planes in ascending order — the whole grid from a context that
// holds it, or one slab rank's planes at a time (out-of-core runs).
struct IcWriter {
    fs::path dir;
    int64_t ppd = 0, chunk = 1, zbeg = 0, zend = 0, last_file = -1;
    int cpd = 1, qdensity = 0, nthreads = 1;
    size_t plane = 0, dplane = 0;
    bool records = true, two = false;
    std::unique_ptr<HostBuffer> buf[2], dbuf[2];
    FILE *densfp   = nullptr;
    WriteStats *ws = nullptr;

    ~IcWriter() { close(); }
    void close() {
        if (densfp) fclose(densfp);
        densfp = nullptr;
    }

    int open(int64_t ppd_, int icformat, const char *output_dir, int cpd_, int qoneslab, WriteStats *ws_, int qdensity_, const char *density_path) {
        ppd = ppd_, cpd = cpd_, ws = ws_, qdensity = qdensity_;
        dir = fs::path(output_dir);
        std::error_code ec;
        if (fs::exists(dir, ec)) {
            for (const auto &entry : fs::directory_iterator(dir, ec)) {
                if (!entry.is_regular_file()) continue;
                const std::string fn = entry.path().filename().string();
                if (fn.compare(0, 3, "ic_") == 0 || fn.compare(0, 10, "zeldovich.") == 0) fs::remove(entry.path(), ec);
            }
        }
        fs::create_directories(dir, ec);
        if (ec) return hfail(ZPLT_EINVAL, "cannot create output directory \"%s\"", output_dir);

        plane  = (size_t) ppd * ppd * zplt_record_bytes(icformat);
        dplane = (size_t) ppd * ppd * sizeof(float);
        size_t chunk_bytes = 1024ull << 20;
        if (const char *e = getenv("ZPLT_IC_CHUNK_BYTES")) chunk_bytes = (size_t) strtoull(e, nullptr, 10);  // tests: small chunks
        chunk = (int64_t) (chunk_bytes / plane);
        if (chunk < 1) chunk = 1;
        if (chunk > ppd) chunk = ppd;
        records = qdensity != 2;
        zbeg = 0, zend = ppd;
        if (qoneslab >= 0) {
            zbeg = qoneslab, zend = qoneslab + 1;
            if (qoneslab >= ppd) zbeg = zend = 0;  // the reference's loop simply never matches
        }
        two = zend - zbeg > chunk;  // a second staging buffer only when there is a second chunk to overlap with
        buf[0].reset(new HostBuffer(records ? (size_t) chunk * plane : 16));
        buf[1].reset(new HostBuffer(records && two ? (size_t) chunk * plane : 16));
        if (!buf[0]->p || !buf[1]->p) return hfail(ZPLT_ENOMEM, "cannot allocate %zu bytes of host staging", (size_t) chunk * plane);
        // ZD_qdensity: float32 density planes appended to one file, opened "wb" (reference src/output.cpp:282-288)
        dbuf[0].reset(new HostBuffer(qdensity ? (size_t) chunk * dplane : 16));
        dbuf[1].reset(new HostBuffer(qdensity && two ? (size_t) chunk * dplane : 16));
        if (!dbuf[0]->p || !dbuf[1]->p) return hfail(ZPLT_ENOMEM, "cannot allocate host staging for the density planes");
        if (qdensity) {
            if (!density_path) return hfail(ZPLT_EINVAL, "ZD_qdensity needs a density file name");
            densfp = fopen(density_path, "wb");
            if (!densfp) return hfail(ZPLT_EINVAL, "cannot open density file \"%s\"", density_path);
        }
        nthreads = (int) std::thread::hardware_concurrency();
        if (nthreads > 8) nthreads = 8;
        if (nthreads < 1) nthreads = 1;
        return ZPLT_OK;
    }

    // the context's planes [0, nplanes) are the global planes [zglobal0, zglobal0 + nplanes)
    int write(zplt_ctx *ctx, int64_t zglobal0, int64_t nplanes) {
        const int64_t a = std::max(zbeg, zglobal0), b = std::min(zend, zglobal0 + nplanes);
        if (a >= b) return ZPLT_OK;
        unsigned char *rbuf[2] = {buf[0]->p, two ? buf[1]->p : buf[0]->p};
        unsigned char *dbf[2]  = {dbuf[0]->p, two ? dbuf[1]->p : dbuf[0]->p};
        auto fetch = [&](int64_t z0, int which) -> int {
            const int64_t nz = (z0 + chunk <= b) ? chunk : b - z0;
            return zplt_fetch_planes_density(ctx, z0 - zglobal0, nz, records ? rbuf[which] : nullptr, qdensity ? (float *) dbf[which] : nullptr);
        };
        int rc    = fetch(a, 0);
        int which = 0;
        for (int64_t z0 = a; rc == ZPLT_OK && z0 < b; z0 += chunk, which ^= 1) {
            const int64_t nz = (z0 + chunk <= b) ? chunk : b - z0;
            // what the writers do with this chunk
            std::vector<WriteJob> jobs;
            for (int64_t z = z0; records && z < z0 + nz; z++) {
                const int64_t fileno = z * cpd / ppd;  // integer division, reference src/output.cpp:208
                if (!jobs.empty() && jobs.back().fileno == fileno) {
                    jobs.back().bytes += plane;
                } else {
                    jobs.push_back({fileno, rbuf[which] + (size_t) (z - z0) * plane, plane});
                    if (fileno != last_file && ws) ws->files++;
                    last_file = fileno;
                }
            }
            const double t0 = now_s();
            std::string werr;
            bool wok    = true;
            int rc_next = ZPLT_OK;
            {
                // the writers work on this chunk while this thread drives the device through the next one
                std::thread writers([&]() {
                    if (densfp && fwrite(dbf[which], 1, (size_t) nz * dplane, densfp) != (size_t) nz * dplane) {
                        wok  = false;
                        werr = "short write on the density file";
                        return;
                    }
                    wok = run_write_jobs(dir, jobs, nthreads, &werr);
                });
                if (z0 + chunk < b) rc_next = fetch(z0 + chunk, which ^ 1);
                writers.join();
            }
            if (ws) {
                ws->seconds += now_s() - t0;
                ws->bytes += (int64_t) ((records ? (size_t) nz * plane : 0) + (densfp ? (size_t) nz * dplane : 0));
            }
            if (!wok)
                rc = hfail(ZPLT_EINVAL, "%s", werr.c_str());
            else
                rc = rc_next;
        }
        return rc;
    }
};

static int write_ic_files(zplt_ctx *ctx, int64_t ppd, int icformat, const char *output_dir, int cpd, int qoneslab, WriteStats *ws,
                          int qdensity = 0, const char *density_path = nullptr) {
    IcWriter w;
    int rc = w.open(ppd, icformat, output_dir, cpd, qoneslab, ws, qdensity, density_path);
    if (rc == ZPLT_OK) rc = w.write(ctx, 0, ppd);
    w.close();
    return rc;
}

// The context does not expose its config through the ABI; keep what the writer needs here.
extern "C" int zplt_ctx_ppd_(const zplt_ctx *ctx);
extern "C" int zplt_ctx_icformat_(const zplt_ctx *ctx);

extern "C" int zplt_write_ic_files(zplt_ctx *ctx, const char *output_dir, int32_t cpd) {
    if (!ctx || !output_dir) return hfail(ZPLT_EINVAL, "null argument");
    return write_ic_files(ctx, zplt_ctx_ppd_(ctx), zplt_ctx_icformat_(ctx), output_dir, cpd, -1, nullptr);
}

extern "C" int zplt_write_outputs(zplt_ctx *ctx, const char *output_dir, int32_t cpd, int32_t qdensity, const char *density_path,
                                  int32_t qoneslab) {
    if (!ctx || !output_dir) return hfail(ZPLT_EINVAL, "null argument");
    return write_ic_files(ctx, zplt_ctx_ppd_(ctx), zplt_ctx_icformat_(ctx), output_dir, cpd, qoneslab, nullptr, qdensity, density_path);
}

// density file name: the reference passes ZD_density_filename through fmt::format with ppd as the only argument
// (reference src/output.cpp:283, default "density{:d}").  The subset of that mini-language a file name can sensibly use:
// "{}", "{0}", "{:d}", "{0:d}" and zero-padded widths such as "{:05d}" become ppd; "{{" and "}}" are literal braces.
std::string zplt_format_density_name(const std::string &pattern, long long ppd) {
    std::string out;
    for (size_t i = 0; i < pattern.size(); i++) {
        const char ch = pattern[i];
        if (ch == '{' && i + 1 < pattern.size() && pattern[i + 1] == '{') {
            out += '{', i++;
        } else if (ch == '}' && i + 1 < pattern.size() && pattern[i + 1] == '}') {
            out += '}', i++;
        } else if (ch == '{') {
            const size_t close = pattern.find('}', i);
            if (close == std::string::npos) {
                out += pattern.substr(i);
                break;
            }
            std::string spec = pattern.substr(i + 1, close - i - 1);  // [index][:[0][width][d]]
            const size_t colon = spec.find(':');
            spec               = colon == std::string::npos ? "" : spec.substr(colon + 1);
            bool zero = !spec.empty() && spec[0] == '0';
            size_t width = 0;
            for (char c : spec)
                if (c >= '0' && c <= '9') width = width * 10 + (size_t) (c - '0');
            std::string num = std::to_string(ppd);
            if (num.size() < width) num = std::string(width - num.size(), zero ? '0' : ' ') + num;
            out += num;
            i = close;
        } else {
            out += ch;
        }
    }
    return out;
}
extern "C" int zplt_format_density_name_(const char *pattern, long long ppd, char *out, size_t cap) {  // test hook
    const std::string r = zplt_format_density_name(pattern ? pattern : "", ppd);
    if (!out || cap == 0) return (int) r.size();
    snprintf(out, cap, "%s", r.c_str());
    return (int) r.size();
}
static std::string density_path_of(const zplt_params &P) {
    return (fs::path(P.output_dir) / zplt_format_density_name(P.density_filename, (long long) P.ppd)).string();
}

// ---------------------------------------------------------------- out of core ----
// The reference compiled with -DDISK never holds the cube: ZeldovichZ leaves it as numblock^2 block files
// TMPDIR/zeldovich.{yblock}/zeldovich.{yblock}.{zblock}, ZeldovichXY reads them back one z-block of planes at a time
// (reference src/block_array.cpp:129-382, src/zeldovich.cpp:891-905).  The same two passes here, with the slab
// decomposition as the blocking: block (s, d) = the rows slab rank s owns on the planes rank d owns = block d of rank s's
// send buffer.  One context plays rank 0..G-1 twice (zplt_slab_set_rank); HBM holds 2/G of the cube.
extern "C" int zplt_copy_d2h_(void *host, const void *dev, size_t bytes);
extern "C" int zplt_copy_h2d_(void *dev, const void *host, size_t bytes);
extern "C" int zplt_device_free_bytes_(int device, size_t *free_b);

extern "C" int zplt_set_device_(int device);
extern "C" int zplt_ctx_device_(const zplt_ctx *ctx);

// Where the blocks wait between the passes.
//   pinned   : page-locked host memory, one copy per block.  Simple, PCIe rate, but page-locking the whole cube costs ~0.5 s
//              per GiB up front (measured: 33 of 44 s at PPD=1024).
//   pageable : ordinary memory that is never page-locked; a few threads move every block in 32 MiB chunks through their own
//              small pinned buffers (device <-> pinned by the copy engine, pinned <-> store by the thread), so the first touch of
//              the store's pages is spread over the threads and overlaps the transfers.
//   disk     : the reference's block files, through one pinned block buffer.
enum StoreKind { STORE_PINNED = 0, STORE_PAGEABLE = 1, STORE_DISK = 2 };

struct BlockStore {
    static constexpr size_t CHUNK = 32u << 20;
    int G      = 0;
    size_t blk = 0;
    bool disk  = false;
    int kind   = STORE_PINNED;
    int device = -1;
    fs::path dir;
    std::vector<std::unique_ptr<HostBuffer>> ram;     // [d]: blocks (0..G-1, d), released once rank d has emitted
    std::unique_ptr<HostBuffer> bounce;                // disk: one block on its way to or from a file
    std::vector<std::unique_ptr<HostBuffer>> lanes;   // pageable: one pinned chunk buffer per copy thread
    double seconds = 0;
    int64_t bytes  = 0;
    bool keep      = false;  // leave the block files where they are (pass 1 of a run split over two invocations)

    // one block between the device and the pageable store, chunk by chunk, every lane a thread
    int move_chunks(unsigned char *host, unsigned char *dev, bool to_host) {
        const size_t nchunks = (blk + CHUNK - 1) / CHUNK;
        std::atomic<size_t> next{0};
        std::atomic<int> rc{ZPLT_OK};
        std::mutex emu;
        std::string emsg;  // the error text is per thread (zplt_last_error): carry a lane's message back to the caller's thread
        auto failed = [&](int code) {
            std::lock_guard<std::mutex> g(emu);
            if (rc.exchange(code) == ZPLT_OK) emsg = zplt_last_error();
        };
        auto lane = [&](int i) {
            if (int r = zplt_set_device_(device)) {
                failed(r);
                return;
            }
            unsigned char *b = lanes[i]->p;
            for (;;) {
                const size_t c = next.fetch_add(1);
                if (c >= nchunks || rc.load() != ZPLT_OK) return;
                const size_t off = c * CHUNK, n = std::min(CHUNK, blk - off);
                int r;
                if (to_host) {
                    if ((r = zplt_copy_d2h_(b, dev + off, n)) == ZPLT_OK) memcpy(host + off, b, n);
                } else {
                    memcpy(b, host + off, n);
                    r = zplt_copy_h2d_(dev + off, b, n);
                }
                if (r != ZPLT_OK) failed(r);
            }
        };
        const int nt = (int) std::min<size_t>(lanes.size(), nchunks);
        std::vector<std::thread> pool;
        for (int i = 1; i < nt; i++) pool.emplace_back(lane, i);
        lane(0);
        for (auto &t : pool) t.join();
        return rc.load() == ZPLT_OK ? ZPLT_OK : hfail(rc.load(), "%s", emsg.c_str());
    }

    ~BlockStore() { close(); }
    fs::path block_dir(int s) const { return dir / ("zeldovich." + std::to_string(s)); }
    fs::path block_file(int s, int d) const { return block_dir(s) / ("zeldovich." + std::to_string(s) + "." + std::to_string(d)); }

    int open(int G_, size_t blk_, int kind_, const fs::path &dir_, int device_) {
        G = G_, blk = blk_, kind = kind_, disk = kind_ == STORE_DISK, dir = dir_, device = device_;
        if (disk) {
            std::error_code ec;
            for (int s = 0; s < G; s++) {
                fs::create_directories(block_dir(s), ec);
                if (ec) return hfail(ZPLT_EINVAL, "cannot create block directory \"%s\"", block_dir(s).c_str());
            }
            bounce.reset(new HostBuffer(blk));
            if (!bounce->p) return hfail(ZPLT_ENOMEM, "cannot allocate a %zu-byte block buffer", blk);
        } else {
            ram.resize(G);
            for (int d = 0; d < G; d++) {
                ram[d].reset(new HostBuffer((size_t) G * blk, kind == STORE_PINNED));
                if (!ram[d]->p) return hfail(ZPLT_ENOMEM, "cannot allocate %zu bytes of host memory for the blocks of pass %d", (size_t) G * blk, d);
            }
            if (kind == STORE_PAGEABLE) {
                int nt = (int) std::thread::hardware_concurrency();
                nt     = std::max(1, std::min(nt, 8));
                for (int i = 0; i < nt; i++) {
                    lanes.emplace_back(new HostBuffer(CHUNK));
                    if (!lanes.back()->p) return hfail(ZPLT_ENOMEM, "cannot allocate the copy threads' buffers");
                }
            }
        }
        return ZPLT_OK;
    }
    int put(int s, int d, const void *dev) {
        const double t0 = now_s();
        int rc;
        if (kind == STORE_PAGEABLE) {
            rc = move_chunks(ram[d]->p + (size_t) s * blk, (unsigned char *) const_cast<void *>(dev), true);
        } else if (!disk) {
            rc = zplt_copy_d2h_(ram[d]->p + (size_t) s * blk, dev, blk);
        } else if ((rc = zplt_copy_d2h_(bounce->p, dev, blk)) == ZPLT_OK) {
            FILE *fp  = fopen(block_file(s, d).c_str(), "wb");
            bool good = fp && fwrite(bounce->p, 1, blk, fp) == blk;
            if (fp) good = (fclose(fp) == 0) && good;
            if (!good) rc = hfail(ZPLT_EINVAL, "cannot write block file \"%s\"", block_file(s, d).c_str());
        }
        seconds += now_s() - t0;
        bytes += (int64_t) blk;
        return rc;
    }
    int get(int s, int d, void *dev) {
        const double t0 = now_s();
        int rc;
        if (kind == STORE_PAGEABLE) {
            rc = move_chunks(ram[d]->p + (size_t) s * blk, (unsigned char *) dev, false);
        } else if (!disk) {
            rc = zplt_copy_h2d_(dev, ram[d]->p + (size_t) s * blk, blk);
        } else {
            FILE *fp  = fopen(block_file(s, d).c_str(), "rb");
            bool good = fp && fread(bounce->p, 1, blk, fp) == blk;
            if (fp) fclose(fp);
            rc = good ? zplt_copy_h2d_(dev, bounce->p, blk) : hfail(ZPLT_EINVAL, "cannot read block file \"%s\"", block_file(s, d).c_str());
            std::error_code ec;
            fs::remove(block_file(s, d), ec);  // the reference's quickdelete: the space goes to the ic files
        }
        seconds += now_s() - t0;
        return rc;
    }
    void release(int d) {
        if (!disk && d < (int) ram.size()) ram[d].reset();
    }
    void close() {
        ram.clear();
        if (disk && !keep) {
            std::error_code ec;
            for (int s = 0; s < G; s++) {
                for (int d = 0; d < G; d++) fs::remove(block_file(s, d), ec);
                fs::remove(block_dir(s), ec);
            }
        }
        bounce.reset();
        lanes.clear();
        G = 0;
    }
};

static int64_t host_mem_available() {  // bytes, 0 if unknown
    std::ifstream f("/proc/meminfo");
    std::string key;
    long long kb;
    while (f >> key >> kb) {
        if (key == "MemAvailable:") return (int64_t) kb * 1024;
        f.ignore(256, '\n');
    }
    return 0;
}

// HBM a context of `G` slab ranks needs on top of its tables: the cube (G = 1), or stage-1 + stage-2 buffer with the room for
// padded planes (csrc/zplt_api.cu, zplt_create), plus the fetch staging and the small areas
static size_t device_need(int64_t N, int na, int G) {
    const size_t cube = (size_t) 16 * na * N * N * N;
    const size_t work = G == 1 ? cube : 2 * (cube / G) + (size_t) (N / G) * 8192 * 16;
    return work + (1ull << 30);
}

// 0 = the cube stays in HBM; otherwise the number of passes.  ZPLT_OOC_PASSES forces a value (tests, or to leave HBM to others).
static int choose_passes(const zplt_params &P, const zplt_config &cfg, int device, int *out) {
    const int64_t N = P.ppd;
    const int na    = cfg.qPLT ? 4 : 2;
    *out            = 0;
    if (const char *e = getenv("ZPLT_OOC_PASSES")) {
        const int G = atoi(e);
        if (G == 0 || G == 1) return ZPLT_OK;
        if (G < 0 || G > 16 || (N / 2) % G) return hfail(ZPLT_EINVAL, "ZPLT_OOC_PASSES=%d: the passes must divide ppd/2=%lld and be at most 16", G, (long long) (N / 2));
        *out = G;
        return ZPLT_OK;
    }
    size_t free_b = 0;
    int rc        = zplt_device_free_bytes_(device, &free_b);
    if (rc) return rc;
    const size_t extra = cfg.f_NL != 0. ? (size_t) 16 * N * N * N : 0;
    if (device_need(N, na, 1) + extra <= free_b) return ZPLT_OK;
    if (cfg.f_NL != 0.)
        return hfail(ZPLT_ENOMEM, "ppd=%lld with ZD_f_NL needs %.1f GB of HBM (%.1f GB free) and does not run out of core; use slab ranks on more GPUs",
                     (long long) N, (device_need(N, na, 1) + extra) / 1e9, free_b / 1e9);
    if (N & (N - 1)) return hfail(ZPLT_ENOMEM, "ppd=%lld needs %.1f GB of HBM (%.1f GB free); only power-of-two grids run out of core", (long long) N, device_need(N, na, 1) / 1e9, free_b / 1e9);
    for (int G = 2; G <= 16; G *= 2) {
        if ((N / 2) % G == 0 && device_need(N, na, G) <= free_b) {
            *out = G;
            return ZPLT_OK;
        }
    }
    return hfail(ZPLT_ENOMEM, "ppd=%lld does not fit this device even in 16 passes (%.1f GB needed, %.1f GB free)", (long long) N,
                 device_need(N, na, 16) / 1e9, free_b / 1e9);
}

// the emission without keeping the records (statistics only)
static int drain_planes(zplt_ctx *ctx, int64_t ppd, int icformat, int64_t nplanes) {
    const size_t plane = (size_t) ppd * ppd * zplt_record_bytes(icformat);
    int64_t chunk      = (int64_t) ((512ull << 20) / plane);
    if (chunk < 1) chunk = 1;
    if (chunk > nplanes) chunk = nplanes;
    HostBuffer buf((size_t) chunk * plane);
    if (!buf.p) return hfail(ZPLT_ENOMEM, "cannot allocate host staging");
    for (int64_t z0 = 0; z0 < nplanes; z0 += chunk) {
        const int64_t nz = (z0 + chunk <= nplanes) ? chunk : nplanes - z0;
        if (int rc = zplt_fetch_planes(ctx, z0, nz, buf.p)) return rc;
    }
    return ZPLT_OK;
}

static int run_out_of_core(zplt_ctx *ctx, const zplt_params &P, const zplt_config &cfg, int G, int write_files, zplt_run_report *rep, WriteStats *ws) {
    void *send = nullptr, *recv = nullptr;
    size_t blk = 0;
    int rc     = zplt_exchange_info(ctx, &send, &recv, &blk);
    if (rc) return rc;
    const int64_t cube_bytes = (int64_t) G * G * (int64_t) blk;
    int kind                 = STORE_PINNED;
    if (const char *e = getenv("ZPLT_OOC_STORE")) {
        if (strcmp(e, "disk") == 0)
            kind = STORE_DISK;
        else if (strcmp(e, "pageable") == 0)
            kind = STORE_PAGEABLE;
        else if (strcmp(e, "ram") != 0)
            return hfail(ZPLT_EINVAL, "ZPLT_OOC_STORE must be \"ram\", \"pageable\" or \"disk\"");
    } else {
        const int64_t avail = host_mem_available();
        if (avail > 0 && cube_bytes + (cube_bytes >> 4) + (8ll << 30) > avail) kind = STORE_DISK;
    }
    // ZPLT_OOC_PART=1: pass 1 only, the block files stay; ZPLT_OOC_PART=2: pass 2 only, from the block files of an earlier
    // invocation with the same parameter file and pass count (the reference's -DPART1 / -DPART2 builds, src/zeldovich.cpp:938-979)
    int part = 0;
    if (const char *e = getenv("ZPLT_OOC_PART")) {
        part = atoi(e);
        if (part < 0 || part > 2) return hfail(ZPLT_EINVAL, "ZPLT_OOC_PART must be 0, 1 or 2");
        if (part && !getenv("ZPLT_OOC_PASSES")) return hfail(ZPLT_EINVAL, "ZPLT_OOC_PART needs ZPLT_OOC_PASSES (both invocations must use the same blocking)");
        if (part) kind = STORE_DISK;
    }
    const bool disk = kind == STORE_DISK;
    fprintf(stderr, "Out of core: %d passes over %.3f GiB, blocks of %.3f GiB buffered %s\n", G, cube_bytes / 1073741824.0,
            blk / 1073741824.0, disk ? "on disk" : (kind == STORE_PAGEABLE ? "in pageable host memory" : "in pinned host memory"));
    IcWriter w;
    if (write_files && part != 1 &&
        (rc = w.open(P.ppd, cfg.icformat, P.output_dir, P.cpd, P.qoneslab, ws, P.qdensity, P.qdensity ? density_path_of(P).c_str() : nullptr)))
        return rc;
    BlockStore st;
    if ((rc = st.open(G, blk, kind, fs::path(P.output_dir), zplt_ctx_device_(ctx)))) return rc;
    st.keep = part == 1;
    // pass 1 (the reference's ZeldovichZ): the rows of rank s — generated, x and z transformed — go out as G blocks
    for (int s = 0; s < G && part != 2; s++) {
        if ((rc = zplt_slab_set_rank(ctx, s))) return rc;
        if ((rc = zplt_generate(ctx))) return rc;
        if ((rc = zplt_synchronize(ctx))) return rc;
        for (int d = 0; d < G; d++)
            if ((rc = st.put(s, d, (const unsigned char *) send + (size_t) d * blk))) return rc;
    }
    // pass 2 (ZeldovichXY): the planes of rank d come back from every source, y transform + records
    const int64_t nloc = P.ppd / G;
    for (int d = 0; d < G && part != 1; d++) {
        if ((rc = zplt_slab_set_rank(ctx, d))) return rc;
        for (int s = 0; s < G; s++)
            if ((rc = st.get(s, d, (unsigned char *) recv + (size_t) s * blk))) return rc;
        st.release(d);
        if ((rc = zplt_exchange_adopt(ctx))) return rc;
        if (write_files)
            rc = w.write(ctx, d * nloc, nloc);
        else
            rc = drain_planes(ctx, P.ppd, cfg.icformat, nloc);
        if (rc) return rc;
    }
    w.close();
    rep->ooc_passes     = G;
    rep->ooc_part       = part;
    rep->ooc_disk       = disk ? 1 : 0;
    rep->ooc_bytes      = st.bytes;
    rep->seconds_blocks = st.seconds;
    st.close();
    return ZPLT_OK;
}

// ---------------------------------------------------------------- whole run -------
extern "C" int zplt_run_param_file(const char *param_file, int32_t device, int32_t write_files, zplt_run_report *rep) {
    if (!param_file) return hfail(ZPLT_EINVAL, "null argument");
    const double t_start = now_s();
    zplt_run_report local;
    if (!rep) rep = &local;
    memset(rep, 0, sizeof(*rep));

    zplt_params P;
    int rc = zplt_params_load(param_file, &P);
    if (rc) return rc;
    zplt_power *pk = nullptr;
    if ((rc = zplt_power_create(&P, &pk))) return rc;
    zplt_config cfg;
    if ((rc = zplt_config_from_params(&P, &cfg))) {
        zplt_power_destroy(pk);
        return rc;
    }
    cfg.device = device;
    int passes = 0;
    if ((rc = choose_passes(P, cfg, device, &passes))) {
        zplt_power_destroy(pk);
        return rc;
    }
    if (passes) {
        if (cfg.f_NL != 0.) rc = hfail(ZPLT_EINVAL, "ZD_f_NL does not run out of core");
        cfg.rank = 0, cfg.nranks = passes;
    }
    zplt_ctx *ctx = nullptr;
    if (rc || (rc = zplt_create(&cfg, &ctx))) {
        zplt_power_destroy(pk);
        return rc;
    }
    auto bail = [&](int code) {
        zplt_destroy(ctx);
        zplt_power_destroy(pk);
        return code;
    };
    if ((rc = zplt_power_apply(pk, ctx))) return bail(rc);
    if (P.qPLT && (rc = zplt_load_eigenmodes_file(ctx, P.PLT_filename))) return bail(rc);
    if (P.k_cutoff != 1)
        fprintf(stderr, "Using k_cutoff = %f (effective ppd = %d)\n", P.k_cutoff, (int) (P.ppd / P.k_cutoff + .5));
    rep->seconds_preamble = now_s() - t_start;
    fprintf(stderr, "Preamble took %f seconds\n", rep->seconds_preamble);

    const double t_dev = now_s();
    WriteStats ws;
    if (passes) {
        if ((rc = run_out_of_core(ctx, P, cfg, passes, write_files, rep, &ws))) return bail(rc);
    } else {
        if ((rc = zplt_generate(ctx))) return bail(rc);
        if (write_files) {
            if ((rc = write_ic_files(ctx, P.ppd, cfg.icformat, P.output_dir, P.cpd, P.qoneslab, &ws, P.qdensity,
                                     P.qdensity ? density_path_of(P).c_str() : nullptr)))
                return bail(rc);
        } else if ((rc = drain_planes(ctx, P.ppd, cfg.icformat, P.ppd))) {  // still run the emission (statistics)
            return bail(rc);
        }
    }
    if ((rc = zplt_synchronize(ctx))) return bail(rc);
    rep->seconds_device = now_s() - t_dev - ws.seconds - rep->seconds_blocks;
    rep->seconds_write  = ws.seconds;
    double tm[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // (pass 2 of a split run never generated in this process: its context has no stage times to read)
    if (rep->ooc_part != 2 && (rc = zplt_get_timings(ctx, tm))) return bail(rc);
    for (int i = 0; i < 4; i++) rep->stage_ms[i] = tm[i];
    if ((rc = zplt_get_stats(ctx, &rep->density_variance, rep->max_disp))) return bail(rc);
    rep->ppd           = P.ppd;
    rep->files_written = ws.files;
    rep->bytes_written = ws.bytes;
    if (rep->ooc_part == 1) {  // nothing emitted yet: no statistics to print (the reference's -DNOPART2 build ends here too)
        fprintf(stderr, "Wrote %d block files (%.2f GB in %.2f sec); run again with ZPLT_OOC_PART=2 for the ic files\n", passes * passes,
                rep->ooc_bytes / 1e9, rep->seconds_blocks);
        rep->seconds_total = now_s() - t_start;
        bail(0);
        return ZPLT_OK;
    }
    rep->rms_density   = sqrt(rep->density_variance / ((double) P.ppd * P.ppd * P.ppd));
    rep->input_sigma   = 0;
    // the stderr summary of reference src/zeldovich.cpp:987-1011
    fprintf(stderr, "The rms density variation of the pixels is %f\n", rep->rms_density);
    rep->sigma_prediction = zplt_power_sigmaR(pk, P.separation / 4.0) * pow(P.boxsize, 1.5);
    fprintf(stderr, "This could be compared to the P(k) prediction of %f\n", rep->sigma_prediction);
    if (P.qdensity != 2) {  // reference src/zeldovich.cpp:998
        fprintf(stderr, "The maximum component-wise displacements are (%g, %g, %g), same units as BoxSize.\n", rep->max_disp[0],
                rep->max_disp[1], rep->max_disp[2]);
        fprintf(stderr,
                "For Abacus' 2LPT implementation to work (assuming FINISH_WAIT_RADIUS = 1),\n\tthis implies a maximum CPD of %d\n",
                (int) (P.boxsize / (2 * fabs(rep->max_disp[2]))));
    }
    fprintf(stderr, "Device stages: generate %.3f ms, z-FFT %.3f ms, y-FFT %.3f ms, x-FFT+emit %.3f ms\n", tm[0], tm[1], tm[2], tm[3]);
    if (passes)
        fprintf(stderr, "Block IO took %.2f sec to move %.2f GB out and back ==> %.1f MB/sec\n", rep->seconds_blocks, 2 * rep->ooc_bytes / 1e9,
                2 * rep->ooc_bytes / 1e6 / (rep->seconds_blocks > 0 ? rep->seconds_blocks : 1e-9));
    if (write_files)
        fprintf(stderr, "Writing ic files took %.3g sec to write %.3g MB ==> %.3g MB/sec\n", ws.seconds, ws.bytes / 1e6,
                ws.bytes / 1e6 / (ws.seconds > 0 ? ws.seconds : 1e-9));
    rep->seconds_total = now_s() - t_start;
    fprintf(stderr, "zeldovich took %.4g sec for ppd %lld ==> %.3g Mpart/sec\n", rep->seconds_total, (long long) P.ppd,
            (double) P.np / 1e6 / rep->seconds_total);
    bail(0);
    return ZPLT_OK;
}
