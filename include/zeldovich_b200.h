/* zeldovich_b200.h — C ABI of the B200-native zeldovich-PLT initial-conditions hot path.
 *
 * The reference (abacusorg/zeldovich-PLT) is a single C++ executable with no FFI; the
 * seam this library replaces is the numerical body of its main():
 *
 *     Setup_FFTW(ppd, ...)                      reference src/zeldovich.cpp:938  (:41-75)
 *     BlockArray array(ppd, numblock, narray…)  reference src/zeldovich.cpp:962-970
 *     ZeldovichZ(array, param, Pk, 0, NULL)     reference src/zeldovich.cpp:971  (:517-601, LoadPlane :278-515)
 *     ZeldovichXY(array, param)                 reference src/zeldovich.cpp:982  (:611-695)
 *       -> WriteParticlesSlab(...)              reference src/output.cpp:41-234
 *     globals eig_vecs / density_variance / max_disp
 *
 * Everything crossing this boundary is a plain pointer, size or POD struct.  All
 * functions return 0 on success or a ZPLT_E* code; zplt_last_error() gives the text.
 * No C++ exception crosses the boundary.  A context is not re-entrant; use one host
 * thread per context (one context per GPU).
 *
 * There is NO CPU fallback: every compute entry point fails with ZPLT_ECUDA when no
 * sm_100-class device is usable.
 */
#ifndef ZELDOVICH_B200_H
#define ZELDOVICH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZPLT_OK 0
#define ZPLT_EINVAL 1   /* bad argument / unsupported configuration */
#define ZPLT_ECUDA 2    /* CUDA runtime failure (including: no device) */
#define ZPLT_ESTATE 3   /* call made in the wrong order */
#define ZPLT_ENOMEM 4

/* ICFormat record layouts, same order as the reference's OutputType enum
 * (reference include/output.h:19-49). */
#define ZPLT_FMT_ZELDOVICH 0   /* u16 i,j,k; pad; double displ[3]              32 B */
#define ZPLT_FMT_RVZEL 1       /* u16 i,j,k; pad; float displ[3]; float vel[3] 32 B */
#define ZPLT_FMT_RVDOUBLEZEL 2 /* u16 i,j,k; pad; double displ[3], vel[3]      56 B */
#define ZPLT_FMT_ZELSIMPLE 3   /* float displ[3]                               12 B */

/* What the kernels need from reference class Parameters (include/parameters.h:14-74)
 * after Parameters::setup() (src/parameters.cpp:97-197). */
typedef struct zplt_config {
    int64_t ppd;          /* particles per dimension: a power of two in 16..2048 (2048 needs nranks >= 4: HBM) runs the fused kernels;
                           * any other even ppd in 16..1024 runs the general path (Bluestein transforms, one GPU, no f_NL) */
    double boxsize;       /* BoxSize */
    int64_t seed;         /* ZD_Seed as the reference widens it: (int) sign-extended (src/power_spectrum.cpp:14) */
    double k_cutoff;      /* ZD_k_cutoff >= 1 */
    int32_t corner_modes; /* ZD_CornerModes */
    int32_t qonemode;     /* ZD_qonemode */
    int32_t one_mode[3];  /* ZD_one_mode */
    int32_t qPLT;         /* ZD_qPLT: 4 packed arrays instead of 2 */
    int32_t qPLTrescale;  /* ZD_qPLT_rescale */
    int32_t fixed_power;  /* ZD_qPk_fix_to_mean */
    double PLT_target_z;  /* ZD_PLT_target_z */
    double z_initial;     /* InitialRedshift */
    double f_cluster;     /* ZD_f_cluster */
    int32_t icformat;     /* ZPLT_FMT_* */
    int32_t device;       /* CUDA device ordinal, or -1 for the current device */
    int32_t rank;         /* slab decomposition: this context's rank ... */
    int32_t nranks;       /* ... of nranks (1 = whole problem on this GPU) */
    double f_NL;          /* ZD_f_NL: local primordial non-Gaussianity (0 = Gaussian); on slab ranks see zplt_potential_begin */
    double n_s;           /* ZD_n_s: spectral index of the primordial power (only used with f_NL) */
    double Omega_M;       /* Omega_M at z = 0 (only used with f_NL) */
} zplt_config;

typedef struct zplt_ctx zplt_ctx;

/* ---- lifetime ------------------------------------------------------------------ */
int zplt_create(const zplt_config *cfg, zplt_ctx **out);
void zplt_destroy(zplt_ctx *ctx);
const char *zplt_last_error(void);
/* bytes of one record of the configured ICFormat (reference sizeof_outputtype, src/output.cpp:254-280) */
size_t zplt_record_bytes(int32_t icformat);
/* number of packed complex arrays: 4 with qPLT, else 2 (reference src/zeldovich.cpp:871-876) */
int zplt_narray(const zplt_ctx *ctx);

/* ---- inputs (host pointers; copied) ------------------------------------------- */
/* Spline of ln P against ln k as reference SplineFunction holds it after spline()
 * (include/spline_function.h:105-139): n sorted nodes x, values y, second derivatives
 * y2; plus PowerSpectrum::normalization and Pk_smooth2 after Normalize()
 * (src/power_spectrum.cpp:186-223).  Replaces the per-mode PowerSpectrum::power call
 * (src/power_spectrum.cpp:225-261). */
int zplt_set_power_spline(zplt_ctx *ctx, int32_t n, const double *x, const double *y, const double *y2,
                          double normalization, double Pk_smooth2);
/* Power-law branch of PowerSpectrum::power (src/power_spectrum.cpp:233-236). */
int zplt_set_power_law(zplt_ctx *ctx, double index, double normalization, double Pk_smooth2);
/* ZD_f_NL only: PowerSpectrum::primordial_norm after Normalize() (reference src/power_spectrum.cpp:221-222),
 * i.e. P(kmin) / kmin^n_s with kmin the smallest positive k of the input table — the scalar behind
 * PowerSpectrum::infer_Tk (:268-274) and the M(k, a) factor of src/zeldovich.cpp:377-386.  zplt_power_apply
 * sets it; callers that use zplt_set_power_spline / _law directly must call this too when f_NL != 0. */
int zplt_set_primordial(zplt_ctx *ctx, double primordial_norm);
/* PLT eigenmode table exactly as load_eigmodes reads it (src/zeldovich.cpp:794-830):
 * double[ppd_e][ppd_e][ppd_e/2+1][4].  Required when qPLT != 0. */
int zplt_set_eigenmodes(zplt_ctx *ctx, int32_t ppd_e, const double *table);

/* ---- device resources ---------------------------------------------------------- */
/* Bytes of device workspace the context needs: narray*ppd^3/nranks complex doubles, twice that
 * (send + receive buffer) when nranks > 1. */
size_t zplt_workspace_bytes(const zplt_ctx *ctx);
/* Optional: hand the context a caller-owned device buffer (e.g. a torch tensor's
 * data_ptr) of at least zplt_workspace_bytes(); otherwise it cudaMallocs its own. */
int zplt_set_workspace(zplt_ctx *ctx, void *device_ptr, size_t bytes);
/* Optional: run on a caller-owned cudaStream_t (passed as void*); default is a private stream. */
int zplt_set_stream(zplt_ctx *ctx, void *cuda_stream);

/* ---- the hot path -------------------------------------------------------------- */
/* The 3-D inverse transform is separable, so the axis order is free; it is x, z, y here (the reference does z, then y
 * and x): generation fuses with the contiguous x rows, and the two strided axes keep 128-byte runs.
 *
 * zplt_generate = ZeldovichZ (reference src/zeldovich.cpp:517-601, LoadPlane :278-515) + the row half of Inverse2dFFT:
 * draw the modes, build the packed spectral arrays and inverse-FFT along x in one kernel, then inverse-FFT along z in
 * place.  Leaves the arrays resident in HBM with the y axis still in Fourier space.  Asynchronous on the stream.
 * With f_NL != 0 it first runs the reference's potential pass (main, src/zeldovich.cpp:945-960: ZeldovichZ with
 * gen_phi = 1, ZeldovichXY_Phi :699-790) on one extra ppd^3 complex array owned by the context. */
int zplt_generate(zplt_ctx *ctx);
/* zplt_emit_planes = the column half of Inverse2dFFT (:88-92 as used at :653-658) + WriteParticlesSlab
 * (src/output.cpp:41-234): inverse-FFT along y and emit the records of planes z0 .. z0+nz-1 in (z, y, x) order into
 * `device_out` (nz*ppd*ppd*record_bytes bytes of device memory, or pinned host memory mapped into the device address
 * space).  The arrays are not modified (emission is repeatable).  Accumulates density_variance and max_disp.
 * Asynchronous. */
int zplt_emit_planes(zplt_ctx *ctx, int64_t z0, int64_t nz, void *device_out);
/* Convenience for host callers: emit planes z0..z0+nz-1 and copy them to `host_out`
 * (pageable or pinned), double-buffered through an internal staging ring.  Synchronous. */
int zplt_fetch_planes(zplt_ctx *ctx, int64_t z0, int64_t nz, void *host_out);
/* ---- slab decomposition (nranks > 1): one context per GPU, one exchange ----------
 * zplt_generate() then runs stage 1 (generation, x and z transforms) on this rank's 2*ppd/(2*nranks)
 * rows and leaves them in the send buffer as nranks contiguous blocks of `bytes_per_peer`
 * bytes, block r = the planes rank r owns.  The CALLER performs the all-to-all (block r of
 * every rank to rank r, received in rank order into `recv`: NCCL all_to_all_single or peer
 * copies) — the y<->z transpose the reference does through BlockArray::StoreBlock/LoadBlock
 * (reference src/block_array.cpp:387-414, 466-504) — and then calls zplt_exchange_done().
 * After that zplt_emit_planes / zplt_fetch_planes address this rank's ppd/nranks planes by
 * LOCAL index (global z = rank*ppd/nranks + local); particle ids carry the global index. */
int zplt_exchange_info(zplt_ctx *ctx, void **send, void **recv, size_t *bytes_per_peer);
int zplt_exchange_done(zplt_ctx *ctx);
/* ---- out of core: one context, the ranks of a slab decomposition one after the other -------
 * The reference built with -DDISK keeps the cube as numblock^2 block files [yblock][zblock] and passes over it twice
 * (reference src/block_array.cpp:129-382, README "Out-of-core").  Here block (s, d) — the rows slab rank s owns, on the
 * planes rank d owns — is exactly block d of rank s's send buffer.  zplt_slab_set_rank() makes the context rank `rank` of
 * cfg.nranks from now on (any order, any number of times); the caller copies the send blocks out after zplt_generate()
 * (stage 1 of rank s) and, for stage 2 of rank d, puts blocks (0..nranks-1, d) into `recv` in source order and calls
 * zplt_exchange_adopt() instead of zplt_exchange_done().  HBM holds 2/nranks of the cube at any time.  Not with peers
 * mapped, not with ZD_f_NL.  zplt_run_param_file() does all of this by itself when the cube does not fit the device. */
int zplt_slab_set_rank(zplt_ctx *ctx, int32_t rank);
int zplt_exchange_adopt(zplt_ctx *ctx);
/* Fused exchange over NVLink peer memory (one process per GPU on one node).  Every rank exports a
 * 64-byte CUDA IPC handle of its library-owned workspace (zplt_ipc_export; do not call
 * zplt_set_workspace), the caller gathers the handles in rank order (e.g. torch.distributed
 * all_gather) and hands them to zplt_ipc_import.  From then on zplt_generate's z-axis FFT kernel
 * stores its results straight into the owner ranks' receive buffers — no separate all-to-all
 * pass; rows land at their true y, receive buffer = [local z][array][y][x].  Two barriers across ranks belong to
 * every step, both the caller's (torch.distributed / MPI; see distributed.PeerExchange.begin / .exchange):
 *   (1) after zplt_generate: synchronise the stream, barrier, then zplt_exchange_done() — a rank may only read its
 *       planes once every peer has finished writing them;
 *   (2) before the NEXT zplt_generate: synchronise (emission done), barrier — a peer's z pass of the next step writes
 *       into the very buffer this rank is still emitting from (write-after-read). */
int zplt_ipc_export(zplt_ctx *ctx, void *handle64);
int zplt_ipc_import(zplt_ctx *ctx, int32_t nranks, const void *handles);
/* Unmap the peers' buffers again.  Every rank must have done so (barrier) before any rank frees its workspace or
 * destroys its context: freeing memory a peer still has mapped is undefined (CUDA IPC rule). */
int zplt_ipc_close(zplt_ctx *ctx);
/* ZD_f_NL != 0 on slab ranks with mapped peers: the potential pass (reference main, src/zeldovich.cpp:945-960) has two
 * transposes of its own, so every step becomes
 *     zplt_potential_begin      phi_g(k) = D/M on this rank's rows, x and z transforms, results stored into the owners' planes
 *     [synchronise + barrier]
 *     zplt_potential_exchange   y transform, phi_g + f_NL phi_g^2, y and x transforms, rows y < ppd/2 stored back to their owners
 *     [synchronise + barrier]
 *     zplt_generate             z transform of the returned rows, then the ordinary generation reading D = conj(phi) M
 * (distributed.PeerExchange.generate does all of it).  The two potential buffers live behind the slab buffers in the
 * workspace (zplt_workspace_bytes counts them). */
int zplt_potential_begin(zplt_ctx *ctx);
int zplt_potential_exchange(zplt_ctx *ctx);
/* Layout of the decomposition (host mirror of the device index math, for tests and bindings):
 * which rank owns row y in stage 1 and in which of its slots. */
int zplt_slab_owner(int64_t ppd, int32_t nranks, int64_t y, int32_t *rank, int32_t *slot);
/* Offset, in complex elements, of x-row (a, z, y) inside rank `rank`'s send buffer (stage 1) or
 * receive buffer (stage 2, the per-source block layout of the caller-run all-to-all; with the fused exchange the
 * receive buffer is simply [local z][a][y][x]); -1 if that rank does not hold it. */
int64_t zplt_slab_offset(int64_t ppd, int32_t nranks, int32_t narray, int32_t stage, int32_t rank, int32_t a, int64_t z, int64_t y);

/* ZD_qdensity (reference src/output.cpp:196,217-224; src/zeldovich.cpp:871-876): the same calls with an
 * extra float32 density plane per z, `dens = Re A0`, in (z, y, x) order.  Pass a NULL record pointer
 * for a density-only run (ZD_qdensity = 2: the reference then skips the displacement arrays; here the
 * records are simply not stored).  Either pointer, not both, may be NULL. */
int zplt_emit_planes_density(zplt_ctx *ctx, int64_t z0, int64_t nz, void *device_out, float *device_density);
int zplt_fetch_planes_density(zplt_ctx *ctx, int64_t z0, int64_t nz, void *host_out, float *host_density);

/* Reset the statistics accumulated by the emit calls. */
int zplt_reset_stats(zplt_ctx *ctx);
/* density_variance = sum over emitted particles of dens^2; max_disp[j] = signed value of
 * the largest |pos[j]| (reference globals, src/output.cpp:28-30,190-197).  Synchronises. */
int zplt_get_stats(zplt_ctx *ctx, double *density_variance, double max_disp[3]);
/* Block until everything queued on the context's stream has finished. */
int zplt_synchronize(zplt_ctx *ctx);
/* Milliseconds the device spent in each stage of the last generate/emit calls
 * (CUDA events on the context's stream): out[0] = mode generation + x FFT (one fused kernel), out[1] = z FFT (on a
 * slab rank with mapped peers: 0, the z pass + exchange overlaps out[0] on a second stream and is inside it),
 * out[2] = 0 (reserved), out[3] = y FFT + emission (summed over emit calls since the last generate).
 * Also out[4..7] = number of kernel launches in each of those stages. */
int zplt_get_timings(zplt_ctx *ctx, double out[8]);
/* Tuning / diagnostic switches of a context: "zring", "yring", "wide_records", "emit_scratch", "emit_prefetch",
 * "slab_groups", "p2p_ctas", "p2p_resident", "p2p_helper", "dit2048", "dit2048_emit", "slab_ring", "gen_persist" (csrc/zplt_internal.h, struct Tuning).  Their defaults
 * come from the environment variables ZPLT_<NAME>, read once in zplt_create — nothing on the launch path calls getenv. */
int zplt_set_option(zplt_ctx *ctx, const char *name, int32_t value);

/* ---- introspection for parity tests (small sizes; host output buffers) --------- */
/* n raw pcg64 outputs starting (off_hi*2^64+off_lo) draws after seeding with `seed`,
 * computed on the device with the same jump-table composition the mode kernel uses. */
int zplt_dbg_pcg_draws(uint64_t seed, uint64_t off_hi, uint64_t off_lo, int64_t n, uint64_t *host_out);
/* For n modes given as signed integer wavevectors k[3*i..3*i+2] (ky in [0,ppd/2)): the
 * two raw draws each mode consumes (raw[2*i], raw[2*i+1]) and their (0,1] images. */
int zplt_dbg_mode_draws(zplt_ctx *ctx, int64_t n, const int32_t *k, uint64_t *host_raw, double *host_u);
/* P(k) at k = sqrt(m)*fundamental for m = 0..count-1, as the mode kernel sees it. */
int zplt_dbg_power_table(zplt_ctx *ctx, int64_t count, double *host_out);
/* The packed spectral arrays [narray][z][y][x] (complex double) before any FFT, formed by the plain
 * one-thread-per-mode kernel (zplt_dbg_spectral) or by the hot generation kernel with its transform skipped
 * (zplt_dbg_spectral_hot: the run walk of the generator, the zero-row skip and the pencil builder of the product path). */
int zplt_dbg_spectral(zplt_ctx *ctx, double *host_out);
int zplt_dbg_spectral_hot(zplt_ctx *ctx, double *host_out);
/* The raw 64-bit draws the HOT generation kernel consumed: host_raw[((z*ppd/2 + y)*ppd + x)*2 + {0,1}] for the primary
 * half 0 <= y < ppd/2; rows the kernel skips as all-masked stay 0. */
int zplt_dbg_hot_draws(zplt_ctx *ctx, uint64_t *host_raw);
/* Same-device stand-ins for the peers of a slab rank (tests of the fused z pass + exchange on one GPU):
 * recv[r] = device pointer of rank r's receive buffer, NULL = discard that rank's share. */
int zplt_dbg_set_peers(zplt_ctx *ctx, int32_t nranks, void *const *recv);
/* The arrays after zplt_generate (z and y transformed, x not yet). */
int zplt_dbg_after_generate(zplt_ctx *ctx, double *host_out);
/* Unnormalised backward 1-D FFTs of length n on host data laid out [batch][n]
 * (row_mode=1, contiguous pencils) or [n][batch] (row_mode=0, strided pencils), in place. */
int zplt_dbg_fft(int32_t n, int64_t batch, int32_t row_mode, double *host_data);
/* variant 0 = as zplt_dbg_fft (the kernels a default context uses), 1 = the plain one-tile-per-CTA kernels,
 * 2 = the 8-pencil decimation kernel (n = 2048, strided pencils). */
int zplt_dbg_fft_variant(int32_t n, int64_t batch, int32_t row_mode, int32_t variant, double *host_data);


/* ==== host side of the boundary (C++ implementation, no device work) ================
 * Mirrors of reference class Parameters / PowerSpectrum / load_eigmodes / the ic_* file
 * writer, exported with C linkage so that tests and bindings can check the host scalars
 * and drive a whole run. */

/* Parameters after parsing + setup() (reference include/parameters.h:14-74). */
typedef struct zplt_params {
    double boxsize, Pk_scale, separation, fundamental, nyquist, k_cutoff;
    double Pk_norm, Pk_sigma, Pk_sigma_ratio, f_cluster, Pk_smooth, Pk_powerlaw_index;
    double z_initial, PLT_target_z, f_NL, n_s, Omega_M;
    int64_t ppd, np;
    int32_t cpd, numblock, qdensity, qascii, qoneslab, seed, qPk_fix_to_mean, qonemode, one_mode[3];
    int32_t qPLT, qPLTrescale, AllowDirectIO, version, CornerModes;
    char Pk_filename[1024], output_dir[1024], density_filename[1024], PLT_filename[1024], ICFormat[64];
} zplt_params;

/* Parse a ParseHeader-style parameter file and run the reference's validity checks
 * (reference src/parameters.cpp:11-197).  ZPLT_EINVAL + message on any failure. */
int zplt_params_load(const char *param_file, zplt_params *out);
/* ICFormat string -> ZPLT_FMT_* (reference src/output.cpp:256-279), -1 if unknown. */
int zplt_icformat_code(const char *icformat);
/* Fill a zplt_config from parsed parameters (device = -1, rank 0 of 1). */
int zplt_config_from_params(const zplt_params *p, zplt_config *out);

/* Host PowerSpectrum: InitFromFile / InitFromPowerLaw + Normalize
 * (reference src/power_spectrum.cpp:130-223). */
typedef struct zplt_power zplt_power;
int zplt_power_create(const zplt_params *p, zplt_power **out);
void zplt_power_destroy(zplt_power *pk);
int zplt_power_info(const zplt_power *pk, int32_t *n_nodes, double *normalization, double *Pk_smooth2);
int zplt_power_arrays(const zplt_power *pk, double *x, double *y, double *y2);
double zplt_power_eval(zplt_power *pk, double wavenumber);  /* PowerSpectrum::power */
double zplt_power_sigmaR(zplt_power *pk, double R);         /* PowerSpectrum::sigmaR */
double zplt_power_infer_Tk(zplt_power *pk, double wavenumber); /* PowerSpectrum::infer_Tk (src/power_spectrum.cpp:268-274), f_NL */
double zplt_power_primordial_norm(const zplt_power *pk);    /* PowerSpectrum::primordial_norm (src/power_spectrum.cpp:221-222) */
/* Hand the spline (or power law) to a device context: calls zplt_set_power_spline/_law. */
int zplt_power_apply(zplt_power *pk, zplt_ctx *ctx);

/* load_eigmodes (reference src/zeldovich.cpp:794-830): read `int32 ppd_e` + table, check the
 * file size, and pass it to zplt_set_eigenmodes. */
int zplt_load_eigenmodes_file(zplt_ctx *ctx, const char *path);

/* SetupOutputDir + the ZeldovichXY write loop (reference src/output.cpp:236-251, :208-212;
 * src/zeldovich.cpp:667-682): remove stale ic_* / zeldovich.* files, then append plane z to
 * `output_dir/ic_{z*cpd/ppd}` in ascending z.  Needs zplt_generate to have run. */
int zplt_write_ic_files(zplt_ctx *ctx, const char *output_dir, int32_t cpd);
/* The same with the ZD_qdensity options: qdensity 0 = records only, 1 = records + `density_path` (float32 planes,
 * ascending z, opened "wb" as the reference does in InitOutputBuffers, src/output.cpp:282-288), 2 = density only;
 * qoneslab >= 0 writes just that plane (reference src/zeldovich.cpp:669). */
int zplt_write_outputs(zplt_ctx *ctx, const char *output_dir, int32_t cpd, int32_t qdensity, const char *density_path,
                       int32_t qoneslab);

/* What `zeldovich <param_file>` does end to end (reference main, src/zeldovich.cpp:848-1032);
 * the report carries the numbers the reference prints on stderr. */
typedef struct zplt_run_report {
    double density_variance, rms_density, max_disp[3];
    double input_sigma, sigma_prediction;
    double seconds_total, seconds_preamble, seconds_device, seconds_write;
    double stage_ms[4];
    int64_t ppd, files_written, bytes_written;
    /* out-of-core runs: passes over the cube (0 = the cube was resident), 1 if the blocks went through files under
     * InitialConditionsDirectory instead of host memory, bytes parked, seconds spent copying/writing/reading them */
    int64_t ooc_passes, ooc_disk, ooc_bytes, ooc_part;
    double seconds_blocks;
} zplt_run_report;
/* The cube is kept in HBM when it fits.  Otherwise — or when the environment variable ZPLT_OOC_PASSES=G forces it — the run
 * goes out of core in G passes (see zplt_slab_set_rank), the blocks parked in host memory, or in files
 * `zeldovich.{s}/zeldovich.{s}.{d}` under InitialConditionsDirectory (the reference's names, src/block_array.cpp:136) when
 * host memory is too small or ZPLT_OOC_STORE=disk.  With block files the two passes can be two invocations, as with the
 * reference's -DPART1 / -DPART2 builds: ZPLT_OOC_PART=1 stops after pass 1 and leaves the block files, ZPLT_OOC_PART=2 (same
 * parameter file, same ZPLT_OOC_PASSES) reads them and writes the ic files.  ZPLT_OOC_STORE=ram is the default page-locked store, =pageable ordinary
 * memory filled by several copy threads through small page-locked buffers (no up-front page-locking of the whole cube). */
int zplt_run_param_file(const char *param_file, int32_t device, int32_t write_files, zplt_run_report *report);

#ifdef __cplusplus
}
#endif
#endif /* ZELDOVICH_B200_H */
