// C ABI of libzeldovich_b200 (see include/zeldovich_b200.h): context management, host
// side table construction, kernel sequencing.  No CPU compute fallback exists here: every
// entry point that needs the device fails with ZPLT_ECUDA when CUDA is unusable.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/zeldovich_b200.h"
#include "zplt_internal.h"

using namespace zplt;

static thread_local std::string g_err;

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t _e = (cudaError_t) (call);                                                        \
        if (_e != cudaSuccess) return fail(ZPLT_ECUDA, "%s failed: %s", #call, cudaGetErrorString(_e)); \
    } while (0)

// ---------------------------------------------------------------- host tables -----
namespace zplt {
u128 pcg_seed_state(uint64_t seed) {
    // engine(itype state) : state_(bump(state + increment())) — reference pcg_random.hpp:427-432
    const u128 M = {ZPLT_PCG_MULT_LO, ZPLT_PCG_MULT_HI};
    const u128 C = {ZPLT_PCG_INC_LO, ZPLT_PCG_INC_HI};
    u128 s       = {seed, 0};
    return add128(mul128(add128(s, C), M), C);
}
// The affine map equal to `delta` LCG steps (Brown's O(log delta) jump-ahead, the method
// behind pcg's advance(), reference pcg_random.hpp:664-686), kept as (mult, plus) so the
// device can compose jumps instead of looping.
Affine pcg_jump(unsigned __int128 delta) {
    u128 cm = {ZPLT_PCG_MULT_LO, ZPLT_PCG_MULT_HI};
    u128 cp = {ZPLT_PCG_INC_LO, ZPLT_PCG_INC_HI};
    Affine a;
    a.mult = {1, 0};
    a.plus = {0, 0};
    const u128 one = {1, 0};
    while (delta > 0) {
        if (delta & 1) {
            a.mult = mul128(a.mult, cm);
            a.plus = add128(mul128(a.plus, cm), cp);
        }
        cp = mul128(add128(cm, one), cp);
        cm = mul128(cm, cm);
        delta >>= 1;
    }
    return a;
}
}  // namespace zplt

static const long long MAXPPD = 65536;  // reference include/zeldovich.h:34

// ---------------------------------------------------------------- context ---------
#define ZPLT_MAX_EMIT_EVENTS 256

struct zplt_ctx {
    zplt_config cfg;
    int N, na, device;
    GenParams gp;
    SlabGeom sg;
    size_t slab_elems;  // complex elements of one slab buffer
    size_t b2_elems;    // complex elements of the receive buffer (slab_elems + room for padded planes)
    size_t phi_elems;   // slab rank with ZD_f_NL: complex elements of one buffer of the potential pass (else 0)
    int phi_stage;      // slab rank with ZD_f_NL: 0 = nothing, 1 = zplt_potential_begin done, 2 = zplt_potential_exchange done
    bool exchanged;
    bool p2p;               // peers' stage-2 buffers are mapped: the z pass stores straight into them
    bool dbg_peers;         // ... through zplt_dbg_set_peers (same-device buffers, tests) rather than CUDA IPC
    cplx *peer_recv[16];
    void *peer_base[16];    // what cudaIpcOpenMemHandle returned (to close)
    double vnorm;
    cudaStream_t stream, copy_stream, xchg_stream;  // xchg_stream: high priority, runs the NVLink-bound z pass of slab groups
    cudaEvent_t ev_group[16], ev_join;
    bool own_stream;
    // device memory
    cplx *cube;
    bool own_cube;
    size_t cube_bytes;
    double *ptab;
    long long ptab_count;
    double *spx, *spy, *spy2;
    int sp_n;          // nodes the spline arrays were allocated for
    size_t eig_bytes;  // bytes the eigenmode table was allocated for
    u128 *ystate;
    Affine *zjump, *xjump;
    double *eig;
    cplx *tw;
    double *stats;
    void *scratch;  // per-SM parking space of the emission kernel (256 SMs x 16 x 512 x 24 B = 50 MB)
    bool have_power, have_eig, generated;
    // ZD_f_NL: the potential (ppd^3 complex), the M(k) table, PowerSpectrum::primordial_norm
    cplx *phi;
    double *mtab;
    double primordial_norm;
    bool have_primordial;
    // fetch staging
    unsigned char *stage_dev[2];
    size_t stage_bytes;
    cudaEvent_t stage_free[2], stage_full[2];
    // timing
    cudaEvent_t ev_gen[4];
    cudaEvent_t ev_emit[2 * ZPLT_MAX_EMIT_EVENTS];
    int n_emit_ev;
    int launches[4];
    // ppd not a power of two: Bluestein length, chirp, transformed conjugate chirp, work array [M][Qb]
    int blu_M;
    cplx *blu_w, *blu_B, *blu_W;
    long long blu_Qb;
    unsigned int *group_flags;  // [32]: [j] set when the generation kernel of row group j has completed (resident z pass), [31] = time-out marker
    Tuning tn;     // switches: environment defaults read once in zplt_create, zplt_set_option afterwards
    LaunchRes lr;  // work counters of the persistent kernels, SM count
};

// ZPLT_<NAME> environment defaults of the tuning switches, read once per context
static void tuning_from_env(Tuning &t) {
    struct {
        const char *name;
        int *v;
    } tab[] = {{"ZPLT_ZRING", &t.zring},           {"ZPLT_YRING", &t.yring},           {"ZPLT_WIDE_RECORDS", &t.wide_records},
               {"ZPLT_EMIT_SCRATCH", &t.emit_scratch}, {"ZPLT_EMIT_PREFETCH", &t.emit_prefetch}, {"ZPLT_SLAB_GROUPS", &t.slab_groups},
               {"ZPLT_P2P_CTAS", &t.p2p_ctas},     {"ZPLT_DIT2048", &t.dit2048},       {"ZPLT_DIT2048_EMIT", &t.dit2048_emit},       {"ZPLT_SLAB_RING", &t.slab_ring}, {"ZPLT_P2P_RESIDENT", &t.p2p_resident}, {"ZPLT_P2P_HELPER", &t.p2p_helper}, {"ZPLT_B2_LAYOUT", &t.b2_layout}, {"ZPLT_B2_PAD", &t.b2_pad},
               {"ZPLT_GEN_PERSIST", &t.gen_persist}};
    for (auto &e : tab) {
        const char *s = getenv(e.name);
        if (s && *s) *e.v = atoi(s);
    }
}

extern "C" int zplt_set_option(zplt_ctx *c, const char *name, int32_t value) {
    if (!c || !name) return fail(ZPLT_EINVAL, "null argument");
    Tuning &t = c->tn;
    struct {
        const char *name;
        int *v;
    } tab[] = {{"zring", &t.zring},           {"yring", &t.yring},           {"wide_records", &t.wide_records},
               {"emit_scratch", &t.emit_scratch}, {"emit_prefetch", &t.emit_prefetch}, {"slab_groups", &t.slab_groups},
               {"p2p_ctas", &t.p2p_ctas},     {"dit2048", &t.dit2048},       {"dit2048_emit", &t.dit2048_emit},       {"slab_ring", &t.slab_ring}, {"p2p_resident", &t.p2p_resident}, {"p2p_helper", &t.p2p_helper}, {"b2_layout", &t.b2_layout}, {"b2_pad", &t.b2_pad},
               {"gen_persist", &t.gen_persist}};
    for (auto &e : tab)
        if (!strcmp(e.name, name)) {
            *e.v = value;
            return ZPLT_OK;
        }
    return fail(ZPLT_EINVAL, "unknown option \"%s\"", name);
}

extern "C" const char *zplt_last_error(void) { return g_err.c_str(); }

extern "C" size_t zplt_record_bytes(int32_t f) {
    switch (f) {
        case ZPLT_FMT_ZELDOVICH: return 32;
        case ZPLT_FMT_RVZEL: return 32;
        case ZPLT_FMT_RVDOUBLEZEL: return 56;
        case ZPLT_FMT_ZELSIMPLE: return 12;
    }
    return 0;
}

extern "C" int zplt_narray(const zplt_ctx *ctx) { return ctx ? ctx->na : 0; }

static int upload(void **dptr, const void *h, size_t bytes, cudaStream_t st) {
    CK(cudaMalloc(dptr, bytes));
    CK(cudaMemcpyAsync(*dptr, h, bytes, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    return 0;
}

static int create_impl(const zplt_config *cfg, zplt_ctx **out, zplt_ctx **partial);

extern "C" int zplt_create(const zplt_config *cfg, zplt_ctx **out) {
    if (!cfg || !out) return fail(ZPLT_EINVAL, "null argument");
    *out             = nullptr;
    zplt_ctx *partial = nullptr;
    int rc            = create_impl(cfg, out, &partial);
    if (rc != ZPLT_OK && partial) {  // release whatever had been set up before the failure
        std::string keep = g_err;
        zplt_destroy(partial);
        g_err = keep;
    }
    return rc;
}

static int create_impl(const zplt_config *cfg, zplt_ctx **out, zplt_ctx **partial) {
    const long long N = cfg->ppd;
    // powers of two: the fused kernels; any other even ppd (the reference's requirement, src/block_array.cpp:38-40) up to 1024:
    // the general path of zplt_generic_kernels.cu (one GPU, no f_NL)
    const bool pow2 = (N & (N - 1)) == 0;
    if (N < 16 || N > 2048 || (N & 1) || (!pow2 && N > 1024))
        return fail(ZPLT_EINVAL, "ppd=%lld unsupported: this build handles even ppd in [16, 1024] and ppd = 2048", N);
    if (!pow2 && (cfg->nranks > 1 || cfg->f_NL != 0.))
        return fail(ZPLT_EINVAL, "ppd=%lld is not a power of two: such grids run on one GPU and without ZD_f_NL", N);
    if (cfg->nranks > 16) return fail(ZPLT_EINVAL, "nranks=%d unsupported: at most 16 ranks (one node)", cfg->nranks);
    if (!(cfg->boxsize > 0)) return fail(ZPLT_EINVAL, "BoxSize must be positive");
    if (!(cfg->k_cutoff >= 1)) return fail(ZPLT_EINVAL, "ZD_k_cutoff must be >= 1");
    if (!(cfg->f_cluster > 0. && cfg->f_cluster <= 1.)) return fail(ZPLT_EINVAL, "ZD_f_cluster must be in (0,1]");
    if (zplt_record_bytes(cfg->icformat) == 0) return fail(ZPLT_EINVAL, "unknown ICFormat code %d", cfg->icformat);
    if (cfg->qPLT && !(cfg->icformat == ZPLT_FMT_RVZEL || cfg->icformat == ZPLT_FMT_RVDOUBLEZEL))
        return fail(ZPLT_EINVAL, "ZD_qPLT requires an RV* ICFormat (reference src/parameters.cpp:169)");
    if (cfg->nranks < 1 || cfg->rank < 0 || cfg->rank >= cfg->nranks) return fail(ZPLT_EINVAL, "bad rank %d of %d", cfg->rank, cfg->nranks);
    if (cfg->nranks > 1 && ((N / 2) % cfg->nranks || N / (2 * cfg->nranks) < 1))
        return fail(ZPLT_EINVAL, "nranks=%d must divide ppd/2=%lld", cfg->nranks, N / 2);
    if (cfg->f_NL != 0. && !(cfg->Omega_M > 0.)) return fail(ZPLT_EINVAL, "ZD_f_NL != 0 needs Omega_M > 0");

    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) return fail(ZPLT_ECUDA, "no CUDA device");
    int dev = cfg->device;
    if (dev < 0) CK(cudaGetDevice(&dev));
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return fail(ZPLT_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);

    zplt_ctx *c = new zplt_ctx();
    memset(c, 0, sizeof(*c));
    *partial  = c;
    c->cfg    = *cfg;
    c->N      = (int) N;
    c->na     = cfg->qPLT ? 4 : 2;
    c->device = dev;
    c->tn     = Tuning();
    tuning_from_env(c->tn);
    c->lr     = LaunchRes();
    c->lr.sms = prop.multiProcessorCount;
    CK(cudaMalloc((void **) &c->lr.counters, 64 * sizeof(unsigned int)));
    CK(cudaMalloc((void **) &c->group_flags, 32 * sizeof(unsigned int)));
    CK(cudaMemset(c->group_flags, 0, 32 * sizeof(unsigned int)));
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    {
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&c->xchg_stream, cudaStreamNonBlocking, hi));
        for (int i = 0; i < 16; i++) CK(cudaEventCreateWithFlags(&c->ev_group[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    }
    c->own_stream = true;
    for (int i = 0; i < 4; i++) CK(cudaEventCreate(&c->ev_gen[i]));
    for (int i = 0; i < 2 * ZPLT_MAX_EMIT_EVENTS; i++) CK(cudaEventCreate(&c->ev_emit[i]));
    for (int i = 0; i < 2; i++) {
        CK(cudaEventCreateWithFlags(&c->stage_free[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->stage_full[i], cudaEventDisableTiming));
    }
    c->sg.N = (int) N, c->sg.G = cfg->nranks, c->sg.rank = cfg->rank, c->sg.h = (int) (N / (2 * cfg->nranks)), c->sg.na = c->na;
    c->sg.log2h = 0;
    while ((1 << c->sg.log2h) < c->sg.h) c->sg.log2h++;
    c->sg.log2G = 0;
    while ((1 << c->sg.log2G) < c->sg.G) c->sg.log2G++;
    c->sg.ly0 = 0, c->sg.nly = c->sg.h;
    c->slab_elems = (size_t) c->na * N * N * N / cfg->nranks;
    // one buffer on a single GPU; stage-1 + stage-2 buffers when the grid is slab-decomposed, followed — with ZD_f_NL — by the
    // two buffers of the potential pass (this rank's rows [z][slot][x], this rank's planes [zl][y][x]), inside the same
    // allocation so that one IPC handle maps everything a peer stores into
    c->phi_elems  = (cfg->nranks > 1 && cfg->f_NL != 0.) ? (size_t) N * N * N / cfg->nranks : 0;
    // the receive buffer can be laid out with padded planes (Tuning::b2_pad): room for the largest padding
    c->b2_elems   = cfg->nranks > 1 ? c->slab_elems + (size_t) (N / cfg->nranks) * ZPLT_B2_PAD_MAX : 0;
    c->cube_bytes = (c->slab_elems + c->b2_elems + 2 * c->phi_elems) * sizeof(cplx);
    c->sg.b2_zstride = (long long) c->na * N * N, c->sg.b2_persrc = 0;

    // derived scalars, written exactly as the reference computes them
    GenParams &g  = c->gp;
    g.N           = (int) N;
    g.half        = (int) (N / 2);
    g.na          = c->na;
    g.corner_modes = cfg->corner_modes;
    g.qonemode     = cfg->qonemode;
    for (int i = 0; i < 3; i++) g.one_mode[i] = cfg->one_mode[i];
    g.qPLT        = cfg->qPLT;
    g.qPLTrescale = cfg->qPLTrescale;
    g.fixed_power = cfg->fixed_power;
    g.fundamental  = 2.0 * M_PI / cfg->boxsize;             // src/parameters.cpp:176
    g.fundamental2 = g.fundamental * g.fundamental;         // src/zeldovich.cpp:301
    double separation = cfg->boxsize / N;                   // src/parameters.cpp:174
    double nyquist    = M_PI / separation;                  // src/parameters.cpp:175
    g.k2_cutoff    = nyquist * nyquist / (cfg->k_cutoff * cfg->k_cutoff);  // src/zeldovich.cpp:318-319
    double ik_cutoff = 1.0 / cfg->k_cutoff;                 // src/zeldovich.cpp:302
    g.kmax         = (int) ((double) (N / 2) * ik_cutoff + .5);            // src/zeldovich.cpp:350
    g.f_cluster    = cfg->f_cluster;
    g.target_f     = (sqrt(1. + 24 * cfg->f_cluster) - 1) / 4.;            // src/zeldovich.cpp:305
    double a_NL = 1.0, a0 = 1.0;
    if (cfg->qPLTrescale) {                                  // src/zeldovich.cpp:307-312
        a_NL = 1. / (1 + cfg->PLT_target_z);
        a0   = 1. / (1 + cfg->z_initial);
    }
    g.growth_ratio = a_NL / a0;
    g.log_growth_ratio = log(g.growth_ratio);
    c->vnorm       = cfg->qPLT ? 1.0 : (sqrt(1. + 24 * cfg->f_cluster) - 1) * .25;  // src/output.cpp:78-82

    // RNG tables (reference src/power_spectrum.cpp:26-37 per-plane generators, and the
    // nskip bookkeeping of src/zeldovich.cpp:335,341 turned into per-row/per-column jumps)
    {
        std::vector<u128> ys(N / 2);
        u128 s0 = pcg_seed_state((uint64_t) cfg->seed);
        Affine plane = pcg_jump((unsigned __int128) 2 * MAXPPD * MAXPPD);
        ys[0]        = s0;
        for (long long i = 1; i < N / 2; i++) ys[i] = apply(plane, ys[i - 1]);
        std::vector<Affine> zj(N), xj(N);
        for (long long i = 0; i < N; i++) {
            long long k  = i > N / 2 ? i - N : i;
            long long km = k < 0 ? k + MAXPPD : k;
            zj[i]        = pcg_jump((unsigned __int128) 2 * MAXPPD * km);
            xj[i]        = pcg_jump((unsigned __int128) 2 * km);
        }
        int rc;
        if ((rc = upload((void **) &c->ystate, ys.data(), ys.size() * sizeof(u128), c->stream))) return rc;
        if ((rc = upload((void **) &c->zjump, zj.data(), zj.size() * sizeof(Affine), c->stream))) return rc;
        if ((rc = upload((void **) &c->xjump, xj.data(), xj.size() * sizeof(Affine), c->stream))) return rc;
        g.ystate = c->ystate;
        g.zjump  = c->zjump;
        g.xjump  = c->xjump;
    }
    // twiddles W_L^j = exp(+2 pi i j / L), L = the length the FFT kernels run at: N, or the Bluestein length M = 2^m >= 2N-1
    c->blu_M = 0;
    if (!pow2) {
        c->blu_M = 32;
        while (c->blu_M < 2 * N - 1) c->blu_M *= 2;
    }
    {
        const long long L = pow2 ? N : c->blu_M;
        const long double PI = 3.14159265358979323846264338327950288L;
        std::vector<cplx> tw(L);
        for (long long j = 0; j < L; j++) {
            long double ang = 2.0L * PI * (long double) j / (long double) L;
            tw[j]           = make_double2((double) cosl(ang), (double) sinl(ang));
        }
        int rc;
        if ((rc = upload((void **) &c->tw, tw.data(), tw.size() * sizeof(cplx), c->stream))) return rc;
        if (!pow2) {
            // chirp w[n] = exp(+i pi n^2 / N) with the phase reduced exactly (n^2 mod 2N), and Bhat = the length-M backward transform
            // of b[m] = conj(w[|m|]) (|m| < N, wrapped), summed directly in long double: M^2 = 4 M terms at most, once per context
            const long long M = c->blu_M;
            std::vector<cplx> w(N), Bh(M);
            std::vector<long double> br(M, 0.0L), bi(M, 0.0L), tr(M), ti(M);
            for (long long n = 0; n < N; n++) {
                const long double ang = PI * (long double) ((n * n) % (2 * N)) / (long double) N;
                w[n]                  = make_double2((double) cosl(ang), (double) sinl(ang));
                br[n] = cosl(ang), bi[n] = -sinl(ang);
                if (n) br[M - n] = br[n], bi[M - n] = bi[n];
            }
            for (long long j = 0; j < M; j++) {
                const long double ang = 2.0L * PI * (long double) j / (long double) M;
                tr[j] = cosl(ang), ti[j] = sinl(ang);
            }
            for (long long k = 0; k < M; k++) {
                long double sr = 0.0L, si = 0.0L;
                for (long long m = 0; m < M; m++) {
                    if (br[m] == 0.0L && bi[m] == 0.0L) continue;
                    const long long t = (m * k) % M;
                    sr += br[m] * tr[t] - bi[m] * ti[t];
                    si += br[m] * ti[t] + bi[m] * tr[t];
                }
                Bh[k] = make_double2((double) sr, (double) si);
            }
            if ((rc = upload((void **) &c->blu_w, w.data(), w.size() * sizeof(cplx), c->stream))) return rc;
            if ((rc = upload((void **) &c->blu_B, Bh.data(), Bh.size() * sizeof(cplx), c->stream))) return rc;
        }
    }
    c->ptab_count = 3LL * (N / 2) * (N / 2) + 1;
    CK(cudaMalloc((void **) &c->ptab, c->ptab_count * sizeof(double)));
    g.ptab = c->ptab;
    CK(cudaMalloc((void **) &c->stats, ZPLT_STAT_SLOTS * 8 * sizeof(double)));
    CK(cudaMalloc((void **) &c->scratch, ZPLT_SCRATCH_BYTES));
    CK(cudaMemsetAsync(c->stats, 0, ZPLT_STAT_SLOTS * 8 * sizeof(double), c->stream));
    *out     = c;
    *partial = nullptr;
    return ZPLT_OK;
}

extern "C" void zplt_destroy(zplt_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int r = 0; r < 16; r++)
        if (c->peer_base[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
    if (c->own_cube && c->cube) cudaFree(c->cube);
    cudaFree(c->ptab);
    cudaFree(c->spx);
    cudaFree(c->spy);
    cudaFree(c->spy2);
    cudaFree(c->ystate);
    cudaFree(c->zjump);
    cudaFree(c->xjump);
    cudaFree(c->eig);
    cudaFree(c->tw);
    cudaFree(c->stats);
    cudaFree(c->scratch);
    cudaFree(c->phi);
    cudaFree(c->mtab);
    cudaFree(c->lr.counters);
    cudaFree(c->group_flags);
    cudaFree(c->blu_w);
    cudaFree(c->blu_B);
    cudaFree(c->blu_W);
    for (int i = 0; i < 2; i++) {
        if (c->stage_dev[i]) cudaFree(c->stage_dev[i]);
        if (c->stage_free[i]) cudaEventDestroy(c->stage_free[i]);
        if (c->stage_full[i]) cudaEventDestroy(c->stage_full[i]);
    }
    for (int i = 0; i < 4; i++)
        if (c->ev_gen[i]) cudaEventDestroy(c->ev_gen[i]);
    for (int i = 0; i < 2 * ZPLT_MAX_EMIT_EVENTS; i++)
        if (c->ev_emit[i]) cudaEventDestroy(c->ev_emit[i]);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->xchg_stream) cudaStreamDestroy(c->xchg_stream);
    for (int i = 0; i < 16; i++)
        if (c->ev_group[i]) cudaEventDestroy(c->ev_group[i]);
    if (c->ev_join) cudaEventDestroy(c->ev_join);
    delete c;
}

// ---------------------------------------------------------------- inputs ----------
static int build_power_table(zplt_ctx *c, int is_powerlaw, double index, int n, double normalization, double smooth2) {
    CK(launch_power_table(c->ptab, c->ptab_count, c->gp.fundamental2, is_powerlaw, index, n, c->spx, c->spy, c->spy2,
                          normalization, smooth2, c->stream));
    c->have_power = true;
    return ZPLT_OK;
}

extern "C" int zplt_set_power_spline(zplt_ctx *c, int32_t n, const double *x, const double *y, const double *y2,
                                     double normalization, double Pk_smooth2) {
    if (!c || !x || !y || !y2 || n < 2) return fail(ZPLT_EINVAL, "bad spline arguments");
    CK(cudaSetDevice(c->device));
    // the tables are re-uploaded in place when their size has not changed: no cudaFree/cudaMalloc (both synchronise the
    // device) on a context that is fed new inputs every step; the copies are ordered on the context's stream
    if (c->sp_n != n) {
        CK(cudaStreamSynchronize(c->stream));
        cudaFree(c->spx), cudaFree(c->spy), cudaFree(c->spy2);
        c->spx = c->spy = c->spy2 = nullptr;
        c->sp_n = 0;
        CK(cudaMalloc((void **) &c->spx, n * sizeof(double)));
        CK(cudaMalloc((void **) &c->spy, n * sizeof(double)));
        CK(cudaMalloc((void **) &c->spy2, n * sizeof(double)));
        c->sp_n = n;
    }
    CK(cudaMemcpyAsync(c->spx, x, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->spy, y, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->spy2, y2, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    int rc = build_power_table(c, 0, 0.0, n, normalization, Pk_smooth2);
    if (rc) return rc;
    CK(cudaStreamSynchronize(c->stream));  // the host arrays are the caller's: they may change after this returns
    return ZPLT_OK;
}

extern "C" int zplt_set_power_law(zplt_ctx *c, double index, double normalization, double Pk_smooth2) {
    if (!c) return fail(ZPLT_EINVAL, "null context");
    CK(cudaSetDevice(c->device));
    return build_power_table(c, 1, index, 0, normalization, Pk_smooth2);
}

extern "C" int zplt_set_primordial(zplt_ctx *c, double primordial_norm) {
    if (!c) return fail(ZPLT_EINVAL, "null context");
    if (!(primordial_norm > 0.)) return fail(ZPLT_EINVAL, "primordial_norm must be positive");
    c->primordial_norm = primordial_norm;
    c->have_primordial = true;
    return ZPLT_OK;
}

extern "C" int zplt_set_eigenmodes(zplt_ctx *c, int32_t ppd_e, const double *table) {
    if (!c || !table || ppd_e < 2 || (ppd_e & 1)) return fail(ZPLT_EINVAL, "bad eigenmode table");
    CK(cudaSetDevice(c->device));
    size_t bytes = (size_t) ppd_e * ppd_e * (ppd_e / 2 + 1) * 4 * sizeof(double);
    if (c->eig_bytes != bytes) {
        CK(cudaStreamSynchronize(c->stream));
        cudaFree(c->eig);
        c->eig       = nullptr;
        c->eig_bytes = 0;
        CK(cudaMalloc((void **) &c->eig, bytes));
        c->eig_bytes = bytes;
    }
    CK(cudaMemcpyAsync(c->eig, table, bytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));  // the host table is the caller's
    c->gp.eig        = c->eig;
    c->gp.pe         = ppd_e;
    c->gp.eig_direct = (ppd_e % c->N == 0);
    c->gp.eig_scale  = ((double) ppd_e) / c->N;
    c->have_eig      = true;
    return ZPLT_OK;
}

// ---------------------------------------------------------------- resources -------
extern "C" size_t zplt_workspace_bytes(const zplt_ctx *c) { return c ? c->cube_bytes : 0; }

extern "C" int zplt_set_workspace(zplt_ctx *c, void *p, size_t bytes) {
    if (!c || !p) return fail(ZPLT_EINVAL, "null argument");
    if (bytes < c->cube_bytes) return fail(ZPLT_EINVAL, "workspace too small: %zu < %zu", bytes, c->cube_bytes);
    if (((uintptr_t) p) & 255) return fail(ZPLT_EINVAL, "workspace must be 256-byte aligned");
    if (c->p2p && !c->dbg_peers)
        return fail(ZPLT_ESTATE, "the workspace is mapped by the peers (zplt_ipc_export); it cannot be replaced");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));  // nothing queued may still use the old buffer
    CK(cudaStreamSynchronize(c->xchg_stream));
    if (c->own_cube && c->cube) cudaFree(c->cube);
    c->cube      = (cplx *) p;
    c->own_cube  = false;
    c->generated = false;
    return ZPLT_OK;
}

extern "C" int zplt_set_stream(zplt_ctx *c, void *s) {
    if (!c) return fail(ZPLT_EINVAL, "null context");
    CK(cudaStreamSynchronize(c->stream));
    if (c->own_stream) cudaStreamDestroy(c->stream);
    c->stream     = (cudaStream_t) s;
    c->own_stream = false;
    return ZPLT_OK;
}

static TileGeom geom_axis(int N, int na, int axis);

static int ensure_cube(zplt_ctx *c) {
    if (c->cube) return ZPLT_OK;
    cudaError_t e = cudaMalloc((void **) &c->cube, c->cube_bytes);
    if (e != cudaSuccess) {
        c->cube = nullptr;
        return fail(ZPLT_ENOMEM, "cudaMalloc of %zu workspace bytes failed: %s", c->cube_bytes, cudaGetErrorString(e));
    }
    c->own_cube = true;
    return ZPLT_OK;
}

static int ready(zplt_ctx *c) {
    if (!c) return fail(ZPLT_EINVAL, "null context");
    if (!c->have_power) return fail(ZPLT_ESTATE, "power spectrum not set");
    if (c->cfg.qPLT && !c->have_eig) return fail(ZPLT_ESTATE, "qPLT set but no eigenmode table given");
    if (c->cfg.f_NL != 0. && !c->have_primordial) return fail(ZPLT_ESTATE, "ZD_f_NL set but zplt_set_primordial has not been called");
    CK(cudaSetDevice(c->device));
    return ensure_cube(c);
}

// ZD_f_NL: the reference's potential pass (main, src/zeldovich.cpp:945-960).  Leaves in c->phi the backward transform of
// phi_g + f_NL phi_g^2 (normalised by ppd^3) and points the generation kernels at it.
static int run_potential(zplt_ctx *c) {
    const int N = c->N;
    if (!c->phi) {
        cudaError_t e = cudaMalloc((void **) &c->phi, (size_t) N * N * N * sizeof(cplx));
        if (e != cudaSuccess) {
            c->phi = nullptr;
            return fail(ZPLT_ENOMEM, "cudaMalloc of the %zu-byte potential array failed: %s", (size_t) N * N * N * sizeof(cplx),
                        cudaGetErrorString(e));
        }
    }
    if (!c->mtab) CK(cudaMalloc((void **) &c->mtab, c->ptab_count * sizeof(double)));
    CK(launch_mfactor_table(c->mtab, c->ptab, c->ptab_count, c->gp.fundamental2, c->primordial_norm, c->cfg.n_s, c->cfg.z_initial,
                            c->cfg.Omega_M, c->stream));
    GenParams g = c->gp;
    g.phi       = nullptr;
    g.mtab      = c->mtab;
    CK(launch_generate_phi(g, c->sg, c->phi, c->stream));
    const int T = fft_tile_T(N);
    const int axes[3] = {0, 2, 1};
    for (int pass = 0; pass < 2; pass++) {
        for (int i = 0; i < 3; i++) CK(launch_fft_tiles_any(N, T, c->phi, geom_axis(N, 1, axes[i]), c->tw, c->tn, c->lr, c->stream));
        if (pass == 0) CK(launch_fnl_local(c->phi, N, (long long) N * N * N, c->cfg.f_NL, c->stream));
    }
    c->gp.phi  = c->phi;
    c->gp.mtab = c->mtab;
    c->gp.phi_zstride = (long long) N * N, c->gp.phi_yshift = 0;
    return ZPLT_OK;
}

// ---- ZD_f_NL on slab ranks: the same pass with its two transposes (reference ZeldovichZ gen_phi + ZeldovichXY_Phi with
// StoreBlock/LoadBlock and StoreBlockForward/LoadBlockForward, src/zeldovich.cpp:699-790, src/block_array.cpp:305-464) ----
// buffers behind the two slab buffers: P1 = this rank's rows [z][slot][x], P2 = this rank's planes [zl][y][x]
static cplx *phi_p1(zplt_ctx *c) { return c->cube + c->slab_elems + c->b2_elems; }
static cplx *phi_p2(zplt_ctx *c) { return c->cube + c->slab_elems + c->b2_elems + c->phi_elems; }

static TileGeom rows_geom(int N, int T, long long nrows_per_block, long long nblocks, long long block_stride) {
    // contiguous rows of N points: nblocks blocks of nrows_per_block consecutive rows, block_stride elements apart
    TileGeom g;
    g.nstride = 1, g.plo_stride = N, g.phi_stride = 0, g.pa = T;
    g.tstride = (long long) T * N, g.grid_x = (int) (nrows_per_block / T);
    g.ostride = block_stride, g.grid_y = (int) nblocks;
    g.astride = 0, g.grid_z = 1;
    return g;
}

static int slab_potential_ready(zplt_ctx *c) {
    if (!c) return fail(ZPLT_EINVAL, "null context");
    if (c->sg.G == 1 || c->cfg.f_NL == 0.) return fail(ZPLT_EINVAL, "the staged potential pass belongs to slab ranks with ZD_f_NL != 0");
    if (!c->p2p) return fail(ZPLT_ESTATE, "ZD_f_NL on slab ranks needs mapped peers (zplt_ipc_import)");
    return ready(c);
}

// Stage 1 (rows): phi_g(k) = D/M on this rank's rows, x transform, z transform with the results stored into the owners' planes.
extern "C" int zplt_potential_begin(zplt_ctx *c) {
    int rc = slab_potential_ready(c);
    if (rc) return rc;
    const int N = c->N, T = fft_tile_T(N), h = c->sg.h;
    if (!c->mtab) CK(cudaMalloc((void **) &c->mtab, c->ptab_count * sizeof(double)));
    CK(launch_mfactor_table(c->mtab, c->ptab, c->ptab_count, c->gp.fundamental2, c->primordial_norm, c->cfg.n_s, c->cfg.z_initial,
                            c->cfg.Omega_M, c->stream));
    GenParams g = c->gp;
    g.phi       = nullptr;
    g.mtab      = c->mtab;
    cplx *P1    = phi_p1(c);
    CK(launch_generate_phi(g, c->sg, P1, c->stream));
    CK(launch_fft_tiles_any(N, T, P1, rows_geom(N, T, (long long) N * 2 * h, 1, 0), c->tw, c->tn, c->lr, c->stream));
    SlabGeom s1 = c->sg;
    s1.na = 1, s1.ly0 = 0, s1.nly = h;
    s1.b2_zstride = (long long) N * N, s1.b2_persrc = 0;  // the potential's planes: [zl][y][x]
    cplx *peers[16];
    for (int r = 0; r < 16; r++) peers[r] = (r < c->sg.G && c->peer_recv[r]) ? c->peer_recv[r] + c->b2_elems + c->phi_elems : nullptr;
    Tuning tn   = c->tn;
    tn.p2p_ctas = 0;  // nothing runs beside it
    CK(launch_fft_tiles_p2p_any(N, T, P1, s1, peers, c->tw, tn, c->lr, GroupSync{nullptr, c->group_flags + 31, 1, nullptr}, c->stream));
    c->phi_stage = 1;
    return ZPLT_OK;
}

// Stage 2 (planes; after a barrier): y transform, phi_g + f_NL phi_g^2 in configuration space, y and x transforms of the real
// result (its forward transform is the conjugate of its backward transform), rows y < N/2 back to their owners.
extern "C" int zplt_potential_exchange(zplt_ctx *c) {
    int rc = slab_potential_ready(c);
    if (rc) return rc;
    if (c->phi_stage != 1) return fail(ZPLT_ESTATE, "zplt_potential_begin (and a barrier across ranks) comes first");
    const int N = c->N, T = fft_tile_T(N), np = N / c->sg.G;
    cplx *P2 = phi_p2(c);
    TileGeom gy;  // pencils along y of the planes [zl][y][x]
    gy.nstride = N, gy.plo_stride = 1, gy.phi_stride = 0, gy.pa = T, gy.tstride = T, gy.grid_x = N / T;
    gy.ostride = (long long) N * N, gy.grid_y = np, gy.astride = 0, gy.grid_z = 1;
    CK(launch_fft_tiles_any(N, T, P2, gy, c->tw, c->tn, c->lr, c->stream));
    CK(launch_fnl_local(P2, N, (long long) c->phi_elems, c->cfg.f_NL, c->stream));
    CK(launch_fft_tiles_any(N, T, P2, gy, c->tw, c->tn, c->lr, c->stream));
    // x transform of the rows that are read back: y < N/2 of every plane
    CK(launch_fft_tiles_any(N, T, P2, rows_geom(N, T, N / 2, np, (long long) N * N), c->tw, c->tn, c->lr, c->stream));
    cplx *peers[16];
    for (int r = 0; r < 16; r++) peers[r] = (r < c->sg.G && c->peer_recv[r]) ? c->peer_recv[r] + c->b2_elems : nullptr;
    CK(launch_phi_return(P2, c->sg, peers, c->stream));
    c->phi_stage = 2;
    return ZPLT_OK;
}

// Stage 3 (rows; after a barrier, inside zplt_generate): z transform of the primary rows; the generation kernels read them.
static int slab_potential_finish(zplt_ctx *c) {
    if (c->phi_stage != 2)
        return fail(ZPLT_ESTATE, "ZD_f_NL on slab ranks: zplt_potential_begin, barrier, zplt_potential_exchange, barrier come before zplt_generate");
    const int N = c->N, T = fft_tile_T(N), h = c->sg.h;
    cplx *P1 = phi_p1(c);
    TileGeom gz;  // pencils along z of the rows [z][slot][x], primary slots only
    gz.nstride = (long long) 2 * h * N, gz.plo_stride = 1, gz.phi_stride = 0, gz.pa = T, gz.tstride = T, gz.grid_x = N / T;
    gz.ostride = N, gz.grid_y = h, gz.astride = 0, gz.grid_z = 1;
    CK(launch_fft_tiles_any(N, T, P1, gz, c->tw, c->tn, c->lr, c->stream));
    c->gp.phi = P1, c->gp.mtab = c->mtab;
    c->gp.phi_zstride = (long long) 2 * h * N, c->gp.phi_yshift = c->sg.log2G;
    c->phi_stage = 0;
    return ZPLT_OK;
}

// ---------------------------------------------------------------- hot path --------
static TileGeom geom_axis(int N, int na, int axis) {
    // cube layout [a][z][y][x]; axis 2 = z (stride N^2), 1 = y (stride N), 0 = x rows
    TileGeom g;
    const int T = fft_tile_T(N);
    g.astride   = (long long) N * N * N;
    g.grid_z    = na;
    g.pa        = T;
    g.phi_stride = 0;
    if (axis == 2) {
        g.nstride = (long long) N * N, g.ostride = N, g.tstride = T, g.plo_stride = 1;
        g.grid_x = N / T, g.grid_y = N;
    } else if (axis == 1) {
        g.nstride = N, g.ostride = (long long) N * N, g.tstride = T, g.plo_stride = 1;
        g.grid_x = N / T, g.grid_y = N;
    } else {
        g.nstride = 1, g.ostride = (long long) N * N, g.tstride = (long long) T * N, g.plo_stride = N;
        g.grid_x = N / T, g.grid_y = N;
    }
    return g;
}

// with_fft = false: the packed spectral arrays without any transform (introspection).  hot = true forms them with the
// fused generation kernel itself (its transform skipped), otherwise with the plain one-thread-per-mode kernel.
static int run_generate(zplt_ctx *c, bool with_fft, bool hot = true) {
    int rc = ready(c);
    if (rc) return rc;
    c->launches[0] = c->launches[1] = c->launches[2] = c->launches[3] = 0;
    const bool slab = c->sg.G > 1;
    if (slab && !with_fft) return fail(ZPLT_EINVAL, "spectral introspection is single-GPU only");
    CK(cudaEventRecord(c->ev_gen[0], c->stream));
    if (c->cfg.f_NL != 0.) {
        if ((rc = slab ? slab_potential_finish(c) : run_potential(c))) return rc;
        c->launches[0] += 10;
    }
    if (c->blu_M) {
        // ppd not a power of two: plain generation kernel, then every axis (x, z, y as elsewhere) by Bluestein's algorithm
        if (!with_fft && hot) return fail(ZPLT_EINVAL, "the fused generation kernel does not exist for this ppd (not a power of two)");
        CK(launch_generate(c->gp, c->cube, c->stream));
        c->launches[0] += 1;
        CK(cudaEventRecord(c->ev_gen[1], c->stream));
        if (with_fft) {
            if (!c->blu_W) {
                const int T = fft_tile_T(c->blu_M);
                long long Qb = (256ll << 20) / ((long long) c->blu_M * (long long) sizeof(cplx));  // a 256 MB work array
                const long long total = (long long) c->na * c->N * c->N;
                if (Qb > total) Qb = total;
                Qb = (Qb + T - 1) / T * T;
                CK(cudaMalloc((void **) &c->blu_W, (size_t) Qb * c->blu_M * sizeof(cplx)));
                c->blu_Qb = Qb;
            }
            const int axes[3] = {0, 2, 1};
            for (int i = 0; i < 3; i++)
                CK(launch_bluestein_axis(c->cube, c->N, c->blu_M, c->na, axes[i], c->blu_W, c->blu_Qb, c->blu_w, c->blu_B, c->tw, c->tn, c->lr,
                                         c->stream));
            c->launches[1] += 15;
        }
        CK(cudaEventRecord(c->ev_gen[2], c->stream));
        CK(cudaEventRecord(c->ev_gen[3], c->stream));
        c->n_emit_ev = 0;
        c->generated = with_fft;
        c->exchanged = false;
        return ZPLT_OK;
    }
    const int gt = (with_fft || hot) ? gen_xfft_T(c->N, c->na) : 0;
    if (slab && c->p2p) {
        // Slab rank with mapped peers: stage 1 runs in groups of rows.  The z pass of group j (NVLink-bound,
        // on the high-priority exchange stream) overlaps with the generation + x pass of group j+1.
        if (!gt) return fail(ZPLT_EINVAL, "no fused generation kernel for this size");
        int J = c->tn.slab_groups;  // measured on 2 GPUs at PPD=1024 (whole step): 1 group 69.4 ms, 4 groups 54.8, 8 groups 53.5, 16 groups 53.0
        if (J < 1) J = 1;
        if (J > 16) J = 16;
        while (J > 1 && (c->sg.h % J)) J--;
        // The two kernels share the SMs: the z pass (each of its CTAs fills an SM) gets Tuning::p2p_ctas of them, generation the
        // rest; stage 1 ~ max(z SM-time * 148/n, link time, generation SM-time * 148/(148-n)).  p2p_resident: ONE launch of the
        // z pass for all groups, made before the generation kernels so that its CTAs hold their SMs from the start; it consumes
        // group j once flags[j] is set (a stream-ordered memset after the generation kernel of group j).  Launched per group
        // instead (p2p_resident = 0), its CTAs have to win whole SMs back from two-per-SM generation CTAs and mostly run after
        // them: measured on 8 GPUs at PPD=1024, generation done after 8.0 ms, z pass only after 14.7 ms (link time ~11 ms).
        // receive layout of this step (the emission below reads what is chosen here)
        {
            int pad = c->tn.b2_pad < 0 ? 0 : (c->tn.b2_pad > ZPLT_B2_PAD_MAX ? ZPLT_B2_PAD_MAX : c->tn.b2_pad);
            pad &= ~7;  // whole 128-byte rows
            c->sg.b2_persrc  = c->tn.b2_layout < 0 ? (c->sg.G >= 4 && c->N <= 1024) : c->tn.b2_layout == 1;
            c->sg.b2_zstride = (long long) c->na * c->N * c->N + pad;
        }
        SlabGeom sg = c->sg;
        sg.nly      = c->sg.h / J;
        const int T = fft_tile_T(c->N);
        if (c->tn.p2p_ctas < 0) c->tn.p2p_ctas = c->N >= 2048 ? 84 : (c->sg.G == 2 ? 56 : 64);  // sweeps: profiles/r02_sweep_{2,8}gpu.jsonl
        if (c->tn.p2p_resident > 0) {
            sg.ly0 = 0;
            // everything that will run beside the waiting z pass must already be loaded (lazy module loading synchronises):
            // the generation kernel (cube = NULL: prepare only) and the driver's memset
            CK(launch_gen_xfft(c->N, gt, c->gp, sg, nullptr, c->tw, c->tn, c->lr, false, c->stream));
            CK(cudaMemsetAsync(c->group_flags, 0, 16 * sizeof(unsigned int), c->stream));
            // when the pass hands out its tiles through a counter, a second launch of it joins on the SMs the generation kernels
            // leave behind when they are done (stage 1 is SM-bound on few ranks: half of the device would idle otherwise)
            const bool helper = c->tn.p2p_helper > 0 && !(c->N == 2048 && c->tn.dit2048 > 0) && fft_tiles_p2p_shares_tiles(c->N, c->tn);
            unsigned int *ctr = nullptr;
            if (helper) {
                ctr = c->lr.counters + (c->lr.next_counter++ & 63);
                CK(cudaMemsetAsync(ctr, 0, sizeof(unsigned int), c->stream));
            }
            CK(cudaEventRecord(c->ev_group[0], c->stream));
            CK(cudaStreamWaitEvent(c->xchg_stream, c->ev_group[0], 0));
            CK(launch_fft_tiles_p2p_any(c->N, T, c->cube, sg, c->peer_recv, c->tw, c->tn, c->lr,
                                        GroupSync{c->group_flags, c->group_flags + 31, J, ctr}, c->xchg_stream));
            for (int j = 0; j < J; j++) {
                sg.ly0 = j * sg.nly;
                CK(launch_gen_xfft(c->N, gt, c->gp, sg, c->cube, c->tw, c->tn, c->lr, false, c->stream));
                CK(cudaMemsetAsync(c->group_flags + j, 1, sizeof(unsigned int), c->stream));
            }
            if (helper) {
                sg.ly0      = 0;
                Tuning tn   = c->tn;
                tn.p2p_ctas = 0;
                CK(launch_fft_tiles_p2p_any(c->N, T, c->cube, sg, c->peer_recv, c->tw, tn, c->lr, GroupSync{nullptr, c->group_flags + 31, J, ctr},
                                            c->stream));
            }
        } else {
            for (int j = 0; j < J; j++) {
                sg.ly0 = j * sg.nly;
                CK(launch_gen_xfft(c->N, gt, c->gp, sg, c->cube, c->tw, c->tn, c->lr, false, c->stream));
                CK(cudaEventRecord(c->ev_group[j], c->stream));
                CK(cudaStreamWaitEvent(c->xchg_stream, c->ev_group[j], 0));
                Tuning tn = c->tn;
                if (j == J - 1) tn.p2p_ctas = 0;  // the z pass of the last group has no generation left to share with
                CK(launch_fft_tiles_p2p_any(c->N, T, c->cube, sg, c->peer_recv, c->tw, tn, c->lr, GroupSync{nullptr, c->group_flags + 31, 1, nullptr},
                                            c->xchg_stream));
            }
        }
        c->launches[0] = J;
        c->launches[1] = J;
        CK(cudaEventRecord(c->ev_gen[1], c->stream));
        CK(cudaEventRecord(c->ev_join, c->xchg_stream));
        CK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    } else {
        if (gt) {
            // fused: draw the modes and transform the x axis in one kernel
            CK(launch_gen_xfft(c->N, gt, c->gp, c->sg, c->cube, c->tw, c->tn, c->lr, !with_fft, c->stream));
            c->launches[0] += 1;
        } else {
            CK(launch_generate(c->gp, c->cube, c->stream));
            c->launches[0] += 1;
        }
        CK(cudaEventRecord(c->ev_gen[1], c->stream));
        if (with_fft) {
            TileGeom g = geom_axis(c->N, c->na, 2);
            if (slab) {
                // stage-1 buffer B1[z][a][slot][x]: pencils along z for each of the na*2h local rows
                const int T = fft_tile_T(c->N);
                g.astride = 0, g.grid_z = 1;
                g.ostride = c->N, g.grid_y = c->na * 2 * c->sg.h;
                g.tstride = T, g.grid_x = c->N / T, g.pa = T, g.plo_stride = 1, g.phi_stride = 0;
                g.nstride = (long long) c->na * 2 * c->sg.h * c->N;
            }
            CK(launch_fft_tiles_any(c->N, fft_tile_T(c->N), c->cube, g, c->tw, c->tn, c->lr, c->stream));
            c->launches[1] = 1;
        }
    }
    CK(cudaEventRecord(c->ev_gen[2], c->stream));
    // the y axis is transformed inside the emission kernel (zplt_emit_planes)
    CK(cudaEventRecord(c->ev_gen[3], c->stream));
    c->n_emit_ev = 0;
    c->generated = with_fft;
    c->exchanged = false;
    return ZPLT_OK;
}

extern "C" int zplt_generate(zplt_ctx *c) { return run_generate(c, true); }

extern "C" int zplt_emit_planes(zplt_ctx *c, int64_t z0, int64_t nz, void *device_out) {
    return zplt_emit_planes_density(c, z0, nz, device_out, nullptr);
}

extern "C" int zplt_emit_planes_density(zplt_ctx *c, int64_t z0, int64_t nz, void *device_out, float *device_density) {
    if (!c || (!device_out && !device_density)) return fail(ZPLT_EINVAL, "null argument");
    if (!c->generated) return fail(ZPLT_ESTATE, "zplt_generate has not run");
    const int nplanes = c->N / c->sg.G;  // planes this rank owns (all of them on a single GPU)
    if (z0 < 0 || nz <= 0 || z0 + nz > nplanes) return fail(ZPLT_EINVAL, "plane range [%lld,%lld) outside [0,%d)", (long long) z0, (long long) (z0 + nz), nplanes);
    if (c->sg.G > 1 && !c->exchanged) return fail(ZPLT_ESTATE, "the slab exchange has not been run (zplt_exchange_info / zplt_exchange_done)");
    CK(cudaSetDevice(c->device));
    EmitParams ep;
    ep.icformat     = c->cfg.icformat;
    ep.record_bytes = (int) zplt_record_bytes(c->cfg.icformat);
    ep.na           = c->na;
    ep.qPLT         = c->cfg.qPLT;
    ep.vnorm        = c->vnorm;
    ep.z0           = z0;
    ep.out          = (unsigned char *) device_out;
    ep.dens         = device_density;
    ep.stats        = c->stats;
    ep.scratch      = c->tn.emit_scratch ? c->scratch : nullptr;
    ep.wide_records = c->tn.wide_records;  // 256-bit record stores: default (-1) = on in the ring emission kernel
    ep.prefetch     = c->tn.emit_prefetch;
    const long long N2 = (long long) c->N * c->N;
    if (c->sg.G == 1) {  // the cube [a][z][y][x]
        ep.astride = N2 * c->N, ep.zstride = N2, ep.zglobal0 = 0, ep.nzl = c->N;
    } else if (c->p2p && !c->sg.b2_persrc) {  // after the fused exchange: [zl][a][y][x], rows at their true y
        ep.astride = N2, ep.zstride = c->sg.b2_zstride, ep.zglobal0 = (long long) c->sg.rank * nplanes, ep.nzl = nplanes;
    } else {  // after a caller-run all-to-all: per-source blocks B2[src][zl][a][slot][x] (zplt_slab.h)
        ep.astride = 0, ep.zstride = 0, ep.zglobal0 = (long long) c->sg.rank * nplanes, ep.nzl = nplanes;
    }
    bool timed = c->n_emit_ev < ZPLT_MAX_EMIT_EVENTS;
    if (timed) CK(cudaEventRecord(c->ev_emit[2 * c->n_emit_ev], c->stream));
    if (c->blu_M) {  // ppd not a power of two: the cube is fully transformed, emission is its own kernel
        CK(launch_emit_plain(c->cube, c->N, z0, nz, ep, c->stream));
        c->launches[3] += 1;
    } else {
        CK(launch_fft_emit_strided(c->N, fft_tile_T(c->N), c->sg.G > 1 ? c->cube + c->slab_elems : c->cube, c->sg, z0, nz, ep, c->tw,
                                   c->tn, c->lr, c->stream, &c->launches[3]));
    }
    if (timed) {
        CK(cudaEventRecord(c->ev_emit[2 * c->n_emit_ev + 1], c->stream));
        c->n_emit_ev++;
    }
    return ZPLT_OK;
}

extern "C" int zplt_fetch_planes(zplt_ctx *c, int64_t z0, int64_t nz, void *host_out) {
    return zplt_fetch_planes_density(c, z0, nz, host_out, nullptr);
}

extern "C" int zplt_fetch_planes_density(zplt_ctx *c, int64_t z0, int64_t nz, void *host_out, float *host_density) {
    if (!c || (!host_out && !host_density)) return fail(ZPLT_EINVAL, "null argument");
    if (!c->generated) return fail(ZPLT_ESTATE, "zplt_generate has not run");
    if (z0 < 0 || nz <= 0 || z0 + nz > c->N / c->sg.G) return fail(ZPLT_EINVAL, "plane range outside this rank's planes");
    CK(cudaSetDevice(c->device));
    const size_t plane  = (size_t) c->N * c->N * zplt_record_bytes(c->cfg.icformat);
    const size_t dplane = (size_t) c->N * c->N * sizeof(float);
    long long chunk     = (long long) ((256ull << 20) / plane);
    if (chunk < 1) chunk = 1;
    if (chunk > nz) chunk = nz;
    // each staging buffer holds `chunk` planes of records followed by `chunk` planes of density
    if (c->stage_bytes < (size_t) chunk * (plane + dplane)) {
        for (int i = 0; i < 2; i++) {
            if (c->stage_dev[i]) cudaFree(c->stage_dev[i]);
            c->stage_dev[i] = nullptr;
            CK(cudaMalloc((void **) &c->stage_dev[i], (size_t) chunk * (plane + dplane)));
        }
        c->stage_bytes = (size_t) chunk * (plane + dplane);
    }
    int i = 0;
    bool used[2] = {false, false};
    for (long long z = z0; z < z0 + nz; z += chunk, i ^= 1) {
        long long n = (z + chunk <= z0 + nz) ? chunk : (z0 + nz - z);
        if (used[i]) CK(cudaStreamWaitEvent(c->stream, c->stage_free[i], 0));
        float *ddev = host_density ? (float *) (c->stage_dev[i] + (size_t) chunk * plane) : nullptr;
        int rc = zplt_emit_planes_density(c, z, n, host_out ? c->stage_dev[i] : nullptr, ddev);
        if (rc) return rc;
        CK(cudaEventRecord(c->stage_full[i], c->stream));
        CK(cudaStreamWaitEvent(c->copy_stream, c->stage_full[i], 0));
        if (host_out)
            CK(cudaMemcpyAsync((unsigned char *) host_out + (size_t) (z - z0) * plane, c->stage_dev[i], (size_t) n * plane,
                               cudaMemcpyDeviceToHost, c->copy_stream));
        if (host_density)
            CK(cudaMemcpyAsync(host_density + (size_t) (z - z0) * c->N * c->N, ddev, (size_t) n * dplane, cudaMemcpyDeviceToHost,
                               c->copy_stream));
        CK(cudaEventRecord(c->stage_free[i], c->copy_stream));
        used[i] = true;
    }
    CK(cudaStreamSynchronize(c->copy_stream));
    CK(cudaStreamSynchronize(c->stream));
    return ZPLT_OK;
}

extern "C" int zplt_exchange_info(zplt_ctx *c, void **send, void **recv, size_t *bytes_per_peer) {
    if (!c) return fail(ZPLT_EINVAL, "null context");
    if (c->sg.G == 1) return fail(ZPLT_EINVAL, "a single-GPU context has no exchange");
    if (c->p2p) return fail(ZPLT_ESTATE, "the exchange is fused into zplt_generate (peers are mapped); there are no blocks to move");
    int rc = ensure_cube(c);
    if (rc) return rc;
    if (send) *send = c->cube;
    if (recv) *recv = c->cube + c->slab_elems;
    if (bytes_per_peer) *bytes_per_peer = (size_t) slab_block_elems(c->sg) * sizeof(cplx);
    return ZPLT_OK;
}

extern "C" int zplt_ipc_export(zplt_ctx *c, void *handle64) {
    if (!c || !handle64) return fail(ZPLT_EINVAL, "null argument");
    if (c->sg.G == 1) return fail(ZPLT_EINVAL, "a single-GPU context has no exchange");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CK(cudaSetDevice(c->device));
    if (c->cube && !c->own_cube) return fail(ZPLT_ESTATE, "peer exchange needs the library-owned workspace (do not call zplt_set_workspace)");
    int rc = ensure_cube(c);
    if (rc) return rc;
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->cube));
    memcpy(handle64, &h, 64);
    return ZPLT_OK;
}

extern "C" int zplt_ipc_import(zplt_ctx *c, int32_t nranks, const void *handles) {
    if (!c || !handles) return fail(ZPLT_EINVAL, "null argument");
    if (nranks != c->sg.G || nranks > 16) return fail(ZPLT_EINVAL, "expected %d handles (at most 16 ranks)", c->sg.G);
    CK(cudaSetDevice(c->device));
    if (!c->cube || !c->own_cube) return fail(ZPLT_ESTATE, "call zplt_ipc_export first");
    for (int r = 0; r < nranks; r++) {
        if (r == c->sg.rank) {
            c->peer_recv[r] = c->cube + c->slab_elems;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *) handles + 64 * r, 64);
        void *base = nullptr;
        CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer_base[r] = base;
        c->peer_recv[r] = (cplx *) base + c->slab_elems;
    }
    c->p2p = true;
    return ZPLT_OK;
}

extern "C" int zplt_ipc_close(zplt_ctx *c) {
    if (!c) return fail(ZPLT_EINVAL, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaStreamSynchronize(c->xchg_stream));
    for (int r = 0; r < 16; r++) {
        if (c->peer_base[r]) CK(cudaIpcCloseMemHandle(c->peer_base[r]));
        c->peer_base[r] = nullptr;
        c->peer_recv[r] = nullptr;
    }
    c->p2p = c->dbg_peers = false;
    c->generated = false;
    return ZPLT_OK;
}

// Test hook: stage-2 buffers on the SAME device stand in for the peers (one GPU runs fft_tile_p2p_kernel and the grouped,
// overlapped stage 1 of every rank of a slab run).  recv[r] = where rank r's stage-2 buffer is, or NULL to discard that
// rank's share (a single rank of a run whose other ranks' buffers would not fit), this context's own included.
extern "C" int zplt_dbg_set_peers(zplt_ctx *c, int32_t nranks, void *const *recv) {
    if (!c || !recv) return fail(ZPLT_EINVAL, "null argument");
    if (nranks != c->sg.G || nranks > 16 || nranks < 2) return fail(ZPLT_EINVAL, "expected %d receive buffers", c->sg.G);
    if (c->p2p && !c->dbg_peers) return fail(ZPLT_ESTATE, "peers are already mapped through CUDA IPC");
    CK(cudaSetDevice(c->device));
    int rc = ensure_cube(c);
    if (rc) return rc;
    for (int r = 0; r < nranks; r++) c->peer_recv[r] = (cplx *) recv[r];
    c->p2p = c->dbg_peers = true;
    return ZPLT_OK;
}

extern "C" int zplt_exchange_done(zplt_ctx *c) {
    if (!c) return fail(ZPLT_EINVAL, "null context");
    if (!c->generated) return fail(ZPLT_ESTATE, "zplt_generate has not run");
    if (c->p2p) {
        unsigned int err = 0;
        CK(cudaSetDevice(c->device));
        CK(cudaMemcpy(&err, c->group_flags + 31, sizeof(err), cudaMemcpyDeviceToHost));
        if (err) {
            cudaMemset(c->group_flags + 31, 0, sizeof(err));
            return fail(ZPLT_ECUDA, "the z pass timed out waiting for a generation kernel (another context holding every SM of this GPU?)");
        }
    }
    c->exchanged = true;
    return ZPLT_OK;
}

// Out-of-core runs: ONE context plays the ranks of a slab decomposition one after the other, the blocks of the exchange parked
// in host memory or files by the caller in between (host/host_api.cpp, run_out_of_core; reference -DDISK BlockArray,
// src/block_array.cpp:129-382).  Nothing the context holds besides SlabGeom::rank depends on the rank: the generator tables are
// per y-plane of the whole grid, the layout math takes the rank as an argument.
extern "C" int zplt_slab_set_rank(zplt_ctx *c, int32_t rank) {
    if (!c) return fail(ZPLT_EINVAL, "null context");
    if (c->sg.G == 1) return fail(ZPLT_EINVAL, "a single-GPU context has no slab ranks");
    if (rank < 0 || rank >= c->sg.G) return fail(ZPLT_EINVAL, "bad rank %d of %d", rank, c->sg.G);
    if (c->p2p) return fail(ZPLT_ESTATE, "peers are mapped: this context is one fixed rank of a running job");
    if (c->cfg.f_NL != 0.) return fail(ZPLT_EINVAL, "ZD_f_NL: the potential pass is staged per rank, a context cannot change rank");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));  // nothing of the previous rank is still in flight
    c->sg.rank = c->cfg.rank = rank;
    c->generated = c->exchanged = false;
    return ZPLT_OK;
}

// The caller has put this rank's blocks B2[src][zl][a][slot][x] into the receive buffer itself (zplt_exchange_info says where),
// without zplt_generate having run for this rank in this context: emission may start.
extern "C" int zplt_exchange_adopt(zplt_ctx *c) {
    if (!c) return fail(ZPLT_EINVAL, "null context");
    if (c->sg.G == 1) return fail(ZPLT_EINVAL, "a single-GPU context has no exchange");
    if (c->p2p) return fail(ZPLT_ESTATE, "the exchange is fused into zplt_generate (peers are mapped)");
    if (!c->cube) return fail(ZPLT_ESTATE, "no workspace yet (zplt_exchange_info allocates it)");
    int rc = ready(c);
    if (rc) return rc;
    c->n_emit_ev = 0;
    c->generated = c->exchanged = true;
    return ZPLT_OK;
}

extern "C" int zplt_slab_owner(int64_t ppd, int32_t nranks, int64_t y, int32_t *rank, int32_t *slot) {
    if (nranks < 1 || ppd < 2 || (ppd / 2) % nranks || y < 0 || y >= ppd || !rank || !slot) return fail(ZPLT_EINVAL, "bad arguments");
    int r, s;
    slab_owner((int) ppd, nranks, (int) y, r, s);
    *rank = r, *slot = s;
    return ZPLT_OK;
}

extern "C" int64_t zplt_slab_offset(int64_t ppd, int32_t nranks, int32_t narray, int32_t stage, int32_t rank, int32_t a, int64_t z,
                                    int64_t y) {
    SlabGeom g;
    g.N = (int) ppd, g.G = nranks, g.rank = rank, g.h = (int) (ppd / (2 * nranks)), g.na = narray;
    g.log2h = 0;
    while ((1 << g.log2h) < g.h) g.log2h++;
    g.log2G = 0;
    while ((1 << g.log2G) < g.G) g.log2G++;
    g.ly0 = 0, g.nly = g.h;
    g.b2_zstride = 0, g.b2_persrc = 1;
    if (stage == 1) {  // where rank `rank` (the owner of row y) keeps row (a, z, y) before the exchange
        int r, s;
        slab_owner(g.N, g.G, (int) y, r, s);
        if (r != rank) return -1;
        return slab_b1_row(g, a, (int) z, s);
    }
    // stage 2: where rank `rank` (the owner of plane z) finds row (a, z, y) after the exchange
    const int np = g.N / g.G;
    if (z / np != rank) return -1;
    return slab_b2_row(g, a, (int) (z % np), (int) y);
}

extern "C" int zplt_reset_stats(zplt_ctx *c) {
    if (!c) return fail(ZPLT_EINVAL, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaMemsetAsync(c->stats, 0, ZPLT_STAT_SLOTS * 8 * sizeof(double), c->stream));
    return ZPLT_OK;
}

extern "C" int zplt_get_stats(zplt_ctx *c, double *var, double md[3]) {
    if (!c) return fail(ZPLT_EINVAL, "null context");
    CK(cudaSetDevice(c->device));
    double h[ZPLT_STAT_SLOTS * 8];
    CK(cudaMemcpyAsync(h, c->stats, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    double v = 0, mp[3] = {0, 0, 0}, mn[3] = {0, 0, 0};
    for (int s = 0; s < ZPLT_STAT_SLOTS; s++) {
        v += h[8 * s];
        for (int j = 0; j < 3; j++) {
            if (h[8 * s + 1 + j] > mp[j]) mp[j] = h[8 * s + 1 + j];
            if (h[8 * s + 4 + j] > mn[j]) mn[j] = h[8 * s + 4 + j];
        }
    }
    if (var) *var = v;
    // signed value of the largest |pos[j]| (reference src/output.cpp:190-193)
    if (md)
        for (int j = 0; j < 3; j++) md[j] = (mp[j] >= mn[j]) ? mp[j] : -mn[j];
    return ZPLT_OK;
}

extern "C" int zplt_synchronize(zplt_ctx *c) {
    if (!c) return fail(ZPLT_EINVAL, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return ZPLT_OK;
}

extern "C" int zplt_get_timings(zplt_ctx *c, double out[8]) {
    if (!c || !out) return fail(ZPLT_EINVAL, "null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 8; i++) out[i] = 0;
    for (int i = 0; i < 3; i++) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, c->ev_gen[i], c->ev_gen[i + 1]));
        out[i] = ms;
    }
    for (int i = 0; i < c->n_emit_ev; i++) {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, c->ev_emit[2 * i], c->ev_emit[2 * i + 1]));
        out[3] += ms;
    }
    for (int i = 0; i < 4; i++) out[4 + i] = c->launches[i];
    return ZPLT_OK;
}

// ---------------------------------------------------------------- introspection ---
extern "C" int zplt_dbg_pcg_draws(uint64_t seed, uint64_t off_hi, uint64_t off_lo, int64_t n, uint64_t *host_out) {
    if (!host_out || n <= 0) return fail(ZPLT_EINVAL, "bad arguments");
    u128 s0  = pcg_seed_state(seed);
    Affine j = pcg_jump(((unsigned __int128) off_hi << 64) | off_lo);
    u128 *ds;
    Affine *dj;
    uint64_t *dout;
    CK(cudaMalloc((void **) &ds, sizeof(u128)));
    CK(cudaMalloc((void **) &dj, sizeof(Affine)));
    CK(cudaMalloc((void **) &dout, n * sizeof(uint64_t)));
    CK(cudaMemcpy(ds, &s0, sizeof(u128), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dj, &j, sizeof(Affine), cudaMemcpyHostToDevice));
    CK(launch_pcg_draws(ds, dj, n, dout, 0));
    CK(cudaMemcpy(host_out, dout, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    cudaFree(ds), cudaFree(dj), cudaFree(dout);
    return ZPLT_OK;
}

extern "C" int zplt_dbg_mode_draws(zplt_ctx *c, int64_t n, const int32_t *k, uint64_t *host_raw, double *host_u) {
    if (!c || !k || !host_raw || !host_u || n <= 0) return fail(ZPLT_EINVAL, "bad arguments");
    CK(cudaSetDevice(c->device));
    for (int64_t i = 0; i < n; i++) {
        int kx = k[3 * i], ky = k[3 * i + 1], kz = k[3 * i + 2];
        if (ky < 0 || ky >= c->N / 2 || kx < -c->N / 2 || kx > c->N / 2 || kz < -c->N / 2 || kz > c->N / 2)
            return fail(ZPLT_EINVAL, "mode %lld outside the primary half-lattice", (long long) i);
    }
    int *dk;
    uint64_t *dr;
    double *du;
    CK(cudaMalloc((void **) &dk, 3 * n * sizeof(int)));
    CK(cudaMalloc((void **) &dr, 2 * n * sizeof(uint64_t)));
    CK(cudaMalloc((void **) &du, 2 * n * sizeof(double)));
    CK(cudaMemcpy(dk, k, 3 * n * sizeof(int), cudaMemcpyHostToDevice));
    CK(launch_mode_draws(c->gp, n, dk, dr, du, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(host_raw, dr, 2 * n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(host_u, du, 2 * n * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(dk), cudaFree(dr), cudaFree(du);
    return ZPLT_OK;
}

extern "C" int zplt_dbg_power_table(zplt_ctx *c, int64_t count, double *host_out) {
    if (!c || !host_out || count <= 0 || count > c->ptab_count) return fail(ZPLT_EINVAL, "bad arguments");
    if (!c->have_power) return fail(ZPLT_ESTATE, "power spectrum not set");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(host_out, c->ptab, count * sizeof(double), cudaMemcpyDeviceToHost));
    return ZPLT_OK;
}

static int dbg_spectral(zplt_ctx *c, double *host_out, bool hot) {
    if (!host_out) return fail(ZPLT_EINVAL, "null argument");
    int rc = run_generate(c, false, hot);
    if (rc) return rc;
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(host_out, c->cube, c->cube_bytes, cudaMemcpyDeviceToHost));
    return ZPLT_OK;
}
extern "C" int zplt_dbg_spectral(zplt_ctx *c, double *host_out) { return dbg_spectral(c, host_out, false); }
extern "C" int zplt_dbg_spectral_hot(zplt_ctx *c, double *host_out) { return dbg_spectral(c, host_out, true); }

extern "C" int zplt_dbg_hot_draws(zplt_ctx *c, uint64_t *host_raw) {
    if (!c || !host_raw) return fail(ZPLT_EINVAL, "null argument");
    if (c->sg.G > 1) return fail(ZPLT_EINVAL, "single-GPU introspection only");
    if (c->cfg.f_NL != 0.) return fail(ZPLT_EINVAL, "with ZD_f_NL the hot kernel reads the potential instead of drawing");
    CK(cudaSetDevice(c->device));
    const size_t bytes = (size_t) c->N * (c->N / 2) * c->N * 2 * sizeof(uint64_t);
    unsigned long long *d = nullptr;
    CK(cudaMalloc((void **) &d, bytes));
    CK(cudaMemsetAsync(d, 0, bytes, c->stream));
    c->gp.dbg_raw = d;
    int rc        = run_generate(c, false, true);
    c->gp.dbg_raw = nullptr;
    if (rc == ZPLT_OK && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = fail(ZPLT_ECUDA, "generation kernel failed");
    if (rc == ZPLT_OK && cudaMemcpy(host_raw, d, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(ZPLT_ECUDA, "copy failed");
    cudaFree(d);
    return rc;
}

extern "C" int zplt_dbg_after_generate(zplt_ctx *c, double *host_out) {
    if (!c || !host_out) return fail(ZPLT_EINVAL, "null argument");
    if (!c->generated) return fail(ZPLT_ESTATE, "zplt_generate has not run");
    if (c->sg.G > 1) return fail(ZPLT_EINVAL, "single-GPU introspection only");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(host_out, c->cube, c->cube_bytes, cudaMemcpyDeviceToHost));
    return ZPLT_OK;
}

extern "C" int zplt_dbg_fft(int32_t n, int64_t batch, int32_t row_mode, double *host_data) {
    return zplt_dbg_fft_variant(n, batch, row_mode, 0, host_data);
}

// variant 0: the kernels a default context uses (ring-prefetched where they exist); 1: the plain one-tile-per-CTA kernels;
// 2: the 8-pencil decimation kernel (n = 2048, strided pencils)
extern "C" int zplt_dbg_fft_variant(int32_t n, int64_t batch, int32_t row_mode, int32_t variant, double *host_data) {
    if (!host_data) return fail(ZPLT_EINVAL, "null argument");
    Tuning tn;
    LaunchRes lr;
    if (variant == 1) tn.zring = 0, tn.dit2048 = 0;
    if (variant == 2) tn.dit2048 = 1;
    const int T = fft_tile_T(n);
    if (T == 0) return fail(ZPLT_EINVAL, "unsupported FFT length %d", n);
    if (batch <= 0 || batch % T) return fail(ZPLT_EINVAL, "batch must be a positive multiple of %d for n=%d", T, n);
    std::vector<cplx> tw(n);
    for (int j = 0; j < n; j++) {
        long double ang = 2.0L * 3.14159265358979323846264338327950288L * (long double) j / (long double) n;
        tw[j]           = make_double2((double) cosl(ang), (double) sinl(ang));
    }
    cplx *dtw, *d;
    size_t bytes = (size_t) n * batch * sizeof(cplx);
    CK(cudaMalloc((void **) &dtw, n * sizeof(cplx)));
    CK(cudaMalloc((void **) &d, bytes));
    CK(cudaMemcpy(dtw, tw.data(), n * sizeof(cplx), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d, host_data, bytes, cudaMemcpyHostToDevice));
    TileGeom g;
    g.astride = 0, g.ostride = 0, g.grid_y = 1, g.grid_z = 1, g.pa = T, g.phi_stride = 0;
    g.grid_x = (int) (batch / T);
    if (row_mode) {
        g.nstride = 1, g.plo_stride = n, g.tstride = (long long) T * n;
    } else {
        g.nstride = batch, g.plo_stride = 1, g.tstride = T;
    }
    {
        int dev = 0;
        CK(cudaGetDevice(&dev));
        CK(cudaDeviceGetAttribute(&lr.sms, cudaDevAttrMultiProcessorCount, dev));
        CK(cudaMalloc((void **) &lr.counters, 64 * sizeof(unsigned int)));
    }
    CK(launch_fft_tiles_any(n, T, d, g, dtw, tn, lr, 0));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(host_data, d, bytes, cudaMemcpyDeviceToHost));
    cudaFree(d), cudaFree(dtw), cudaFree(lr.counters);
    return ZPLT_OK;
}

// ---------------------------------------------------------------- internal hooks --
// used by host/host_api.cpp (same shared object); not part of the public header
extern "C" void zplt_set_error_(const char *msg) { g_err = msg ? msg : ""; }
extern "C" int zplt_ctx_ppd_(const zplt_ctx *c) { return c ? c->N : 0; }
extern "C" int zplt_ctx_icformat_(const zplt_ctx *c) { return c ? c->cfg.icformat : -1; }
extern "C" void *zplt_pinned_alloc_(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void zplt_pinned_free_(void *p) {
    if (p) cudaFreeHost(p);
}
// plain copies between the workspace and host memory, and the free HBM of a device (out-of-core driver, host/host_api.cpp)
extern "C" int zplt_copy_d2h_(void *host, const void *dev, size_t bytes) {
    CK(cudaMemcpy(host, dev, bytes, cudaMemcpyDeviceToHost));
    return ZPLT_OK;
}
extern "C" int zplt_copy_h2d_(void *dev, const void *host, size_t bytes) {
    CK(cudaMemcpy(dev, host, bytes, cudaMemcpyHostToDevice));
    return ZPLT_OK;
}
extern "C" int zplt_set_device_(int device) {  // a host thread of the out-of-core block store joins the context's device
    if (device >= 0) CK(cudaSetDevice(device));
    return ZPLT_OK;
}
extern "C" int zplt_ctx_device_(const zplt_ctx *c) { return c ? c->device : -1; }
extern "C" int zplt_device_free_bytes_(int device, size_t *free_b) {
    size_t total = 0;
    if (device >= 0) CK(cudaSetDevice(device));
    CK(cudaMemGetInfo(free_b, &total));
    return ZPLT_OK;
}
