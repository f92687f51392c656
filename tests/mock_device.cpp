// TEST INFRASTRUCTURE — a stand-in for the device half of the library (csrc/zplt_api.cu) so that the host half
// (host/host_api.cpp: file writer, out-of-core block store, pass scheduling of zplt_run_param_file) can be driven on a
// machine without a GPU.  It computes nothing: "generation" writes labels into the send blocks, "adoption" checks that the
// blocks that came back are the right ones in the right order, "records" encode their global plane number.  Built and used
// only by tests/test_host.py (together with the unmodified host sources); never part of the product library.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/zeldovich_b200.h"

static thread_local std::string g_err;  // per thread, as in the device half
extern "C" void zplt_set_error_(const char *msg) { g_err = msg ? msg : ""; }
extern "C" const char *zplt_last_error(void) { return g_err.c_str(); }
static int fail(int code, const char *msg) {
    g_err = msg;
    return code;
}

struct zplt_ctx {
    zplt_config cfg;
    int N, na, G, rank;
    size_t blk;  // bytes of one block
    std::vector<unsigned char> send, recv;
    bool power = false, eig = false, generated = false, exchanged = false;
    int64_t planes_emitted = 0, generates = 0;
};

static double label(int s, int d, size_t i) { return s * 1.0e6 + d * 1.0e3 + (double) (i % 997); }

extern "C" size_t zplt_record_bytes(int32_t f) {
    switch (f) {
        case ZPLT_FMT_ZELDOVICH: return 32;
        case ZPLT_FMT_RVZEL: return 32;
        case ZPLT_FMT_RVDOUBLEZEL: return 56;
        case ZPLT_FMT_ZELSIMPLE: return 12;
    }
    return 0;
}
extern "C" int zplt_create(const zplt_config *cfg, zplt_ctx **out) {
    if (!cfg || !out) return fail(ZPLT_EINVAL, "null argument");
    if (cfg->nranks < 1 || cfg->nranks > 16 || (cfg->nranks > 1 && (cfg->ppd / 2) % cfg->nranks)) return fail(ZPLT_EINVAL, "mock: bad nranks");
    zplt_ctx *c = new zplt_ctx();
    c->cfg = *cfg, c->N = (int) cfg->ppd, c->na = cfg->qPLT ? 4 : 2, c->G = cfg->nranks, c->rank = cfg->rank;
    const size_t slab = (size_t) 16 * c->na * c->N * c->N * c->N / c->G;
    c->blk            = slab / c->G;
    if (c->G > 1) c->send.resize(slab), c->recv.resize(slab);
    *out = c;
    return ZPLT_OK;
}
extern "C" void zplt_destroy(zplt_ctx *c) { delete c; }
extern "C" int zplt_set_power_spline(zplt_ctx *c, int32_t, const double *, const double *, const double *, double, double) {
    c->power = true;
    return ZPLT_OK;
}
extern "C" int zplt_set_power_law(zplt_ctx *c, double, double, double) {
    c->power = true;
    return ZPLT_OK;
}
extern "C" int zplt_set_primordial(zplt_ctx *, double) { return ZPLT_OK; }
extern "C" int zplt_set_eigenmodes(zplt_ctx *c, int32_t, const double *) {
    c->eig = true;
    return ZPLT_OK;
}
extern "C" int zplt_generate(zplt_ctx *c) {
    if (!c->power) return fail(ZPLT_ESTATE, "mock: power spectrum not set");
    if (c->cfg.qPLT && !c->eig) return fail(ZPLT_ESTATE, "mock: no eigenmodes");
    if (c->G > 1) {
        double *p      = reinterpret_cast<double *>(c->send.data());
        const size_t n = c->blk / 8;
        for (int d = 0; d < c->G; d++)
            for (size_t i = 0; i < n; i++) p[(size_t) d * n + i] = label(c->rank, d, i);
    }
    c->generates++;
    c->generated = true, c->exchanged = false;
    return ZPLT_OK;
}
extern "C" int zplt_exchange_info(zplt_ctx *c, void **send, void **recv, size_t *bytes_per_peer) {
    if (c->G == 1) return fail(ZPLT_EINVAL, "mock: single-GPU context");
    if (send) *send = c->send.data();
    if (recv) *recv = c->recv.data();
    if (bytes_per_peer) *bytes_per_peer = c->blk;
    return ZPLT_OK;
}
extern "C" int zplt_slab_set_rank(zplt_ctx *c, int32_t rank) {
    if (c->G == 1 || rank < 0 || rank >= c->G) return fail(ZPLT_EINVAL, "mock: bad rank");
    c->rank = c->cfg.rank = rank;
    c->generated = c->exchanged = false;
    memset(c->send.data(), 0xff, c->send.size());  // whatever the previous rank left must not be what is used next
    return ZPLT_OK;
}
extern "C" int zplt_exchange_adopt(zplt_ctx *c) {
    if (c->G == 1) return fail(ZPLT_EINVAL, "mock: single-GPU context");
    const double *p = reinterpret_cast<const double *>(c->recv.data());
    const size_t n  = c->blk / 8;
    for (int s = 0; s < c->G; s++)
        for (size_t i = 0; i < n; i++)
            if (p[(size_t) s * n + i] != label(s, c->rank, i)) {
                char buf[200];
                snprintf(buf, sizeof(buf), "mock: receive block %d of rank %d holds %.1f at %zu, expected %.1f", s, c->rank, p[(size_t) s * n + i], i,
                         label(s, c->rank, i));
                return fail(ZPLT_ESTATE, buf);
            }
    c->generated = c->exchanged = true;
    return ZPLT_OK;
}
extern "C" int zplt_fetch_planes_density(zplt_ctx *c, int64_t z0, int64_t nz, void *host_out, float *host_density) {
    if (!c->generated) return fail(ZPLT_ESTATE, "mock: zplt_generate has not run");
    if (c->G > 1 && !c->exchanged) return fail(ZPLT_ESTATE, "mock: exchange has not run");
    const int64_t nloc = c->N / c->G;
    if (z0 < 0 || nz <= 0 || z0 + nz > nloc) return fail(ZPLT_EINVAL, "mock: plane range outside this rank's planes");
    const size_t words = (size_t) c->N * c->N * zplt_record_bytes(c->cfg.icformat) / 4;
    for (int64_t z = z0; z < z0 + nz; z++) {
        const uint32_t gz = (uint32_t) (c->rank * nloc + z);
        if (host_out) {
            uint32_t *w = reinterpret_cast<uint32_t *>(host_out) + (size_t) (z - z0) * words;
            for (size_t j = 0; j < words; j++) w[j] = (gz << 20) ^ (uint32_t) (j & 0xfffff);
        }
        if (host_density)
            for (size_t j = 0; j < (size_t) c->N * c->N; j++) host_density[(size_t) (z - z0) * c->N * c->N + j] = (float) gz;
    }
    c->planes_emitted += nz;
    return ZPLT_OK;
}
extern "C" int zplt_fetch_planes(zplt_ctx *c, int64_t z0, int64_t nz, void *host_out) { return zplt_fetch_planes_density(c, z0, nz, host_out, nullptr); }
extern "C" int zplt_get_stats(zplt_ctx *c, double *var, double md[3]) {
    if (var) *var = (double) c->planes_emitted;  // lets the test count the planes that went through the emission
    if (md) md[0] = (double) c->generates, md[1] = md[2] = 1.0;
    return ZPLT_OK;
}
extern "C" int zplt_synchronize(zplt_ctx *) { return ZPLT_OK; }
extern "C" int zplt_get_timings(zplt_ctx *, double out[8]) {
    for (int i = 0; i < 8; i++) out[i] = 0;
    return ZPLT_OK;
}
extern "C" int zplt_ctx_ppd_(const zplt_ctx *c) { return c->N; }
extern "C" int zplt_ctx_icformat_(const zplt_ctx *c) { return c->cfg.icformat; }
extern "C" void *zplt_pinned_alloc_(size_t bytes) { return malloc(bytes); }
extern "C" void zplt_pinned_free_(void *p) { free(p); }
// ZPLT_MOCK_FAIL_COPY=n: the n-th copy of the process (counted from 1, either direction) fails like a CUDA error would
#include <atomic>
static std::atomic<long> g_copies{0};
static bool copy_fails() {
    const char *e = getenv("ZPLT_MOCK_FAIL_COPY");
    return e && ++g_copies == atol(e);
}
extern "C" void zplt_mock_reset_copies(void) { g_copies = 0; }
extern "C" int zplt_copy_d2h_(void *host, const void *dev, size_t bytes) {
    if (copy_fails()) return fail(ZPLT_ECUDA, "mock: cudaMemcpy device to host failed");
    memcpy(host, dev, bytes);
    return ZPLT_OK;
}
extern "C" int zplt_copy_h2d_(void *dev, const void *host, size_t bytes) {
    if (copy_fails()) return fail(ZPLT_ECUDA, "mock: cudaMemcpy host to device failed");
    memcpy(dev, host, bytes);
    return ZPLT_OK;
}
extern "C" int zplt_set_device_(int) { return ZPLT_OK; }
extern "C" int zplt_ctx_device_(const zplt_ctx *) { return 0; }
extern "C" int zplt_device_free_bytes_(int, size_t *free_b) {
    const char *e = getenv("ZPLT_MOCK_FREE_BYTES");
    *free_b       = e ? (size_t) strtoull(e, nullptr, 10) : (size_t) 180 << 30;
    return ZPLT_OK;
}
