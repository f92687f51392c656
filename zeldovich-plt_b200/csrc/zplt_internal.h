// Internal launcher prototypes shared by the translation units of libzeldovich_b200.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "zplt_device.cuh"
#include "zplt_slab.h"

namespace zplt {

// Where a tile of pencils lives in memory, all in units of complex elements.
// Pencil p of the tile starts at  base + (p % pa)*plo_stride + (p / pa)*phi_stride,
// its n-th point is nstride further.  base = blockIdx.z*astride + blockIdx.y*ostride
// + blockIdx.x*tstride.
struct TileGeom {
    long long astride, ostride, tstride;
    long long plo_stride, phi_stride, nstride;
    int pa;
    int grid_x, grid_y, grid_z;
};

struct EmitParams {
    int icformat;
    int record_bytes;
    int na;             // arrays (2 or 4)
    int qPLT;
    double vnorm;       // velocity factor when !qPLT (reference src/output.cpp:78-82)
    long long z0;       // first plane of this launch (records are written relative to it)
    unsigned char *out; // records, (z - z0, y, x) order; NULL: no records (ZD_qdensity = 2)
    float *dens;        // optional density planes, float32, (z - z0, y, x) order (ZD_qdensity); NULL: none
    double *stats;      // [ZPLT_STAT_SLOTS][8]: sum dens^2, +max[3], -max[3], pad
    int prefetch;       // L2-prefetch the next packed array of the tile during the transform
    int wide_records;   // RVZel + qPLT: write each 32-byte record with one 256-bit store; -1 = default (ring kernel only), 0 off, 1 on
    void *scratch;      // per-CTA-slot, L2-resident parking space [slot][16][threads] x 24 B (qPLT, one CTA per SM), or NULL
    // where the planes live (complex elements): packed array a of local plane zl starts at a*astride + zl*zstride, rows are N apart.
    // Single GPU: the cube [a][z][y][x] (astride = N^3, zstride = N^2).  Slab rank after the fused exchange: [zl][a][y][x]
    // (astride = N^2, zstride = na*N^2).  zglobal0 = global index of local plane 0 (particle ids carry global z).
    long long astride, zstride, zglobal0;
    int nzl;            // local planes held
};
#define ZPLT_STAT_SLOTS 64
#define ZPLT_B2_PAD_MAX 8192
// EmitParams::scratch: [256 SMs][16][512 threads] x 24 B of record parking, then [256 SMs] x 64 KB that replace the
// shared-memory parking area in the persistent ring emission kernel
#define ZPLT_SCRATCH_PARK_BYTES ((size_t) 256 * 16 * 512 * 24)
#define ZPLT_SCRATCH_BYTES (ZPLT_SCRATCH_PARK_BYTES + (size_t) 256 * 65536)

// Tuning / diagnostic switches of a context.  Defaults come from the environment (ZPLT_<NAME>) ONCE, in zplt_create;
// zplt_set_option changes them afterwards.  Nothing on the launch path calls getenv.
struct Tuning {
    int zring        = 12;  // ZPLT_ZRING: slices of the next tile the ring-prefetched strided pass requests through TMA (0 = plain kernel)
    int yring        = 12;  // ZPLT_YRING: the same for the y pass + emission (0 = one tile per CTA)
    int wide_records = -1;  // ZPLT_WIDE_RECORDS: 256-bit RVZel record stores (qPLT records written whole from the parking areas); 0 = off
    int emit_scratch = 1;   // ZPLT_EMIT_SCRATCH: park record halves in the L2-resident scratch (whole-record stores)
    int emit_prefetch = 1;  // ZPLT_EMIT_PREFETCH: L2-prefetch the next packed array of a tile (one-tile-per-CTA kernel)
    int slab_groups  = 16;  // ZPLT_SLAB_GROUPS: row groups of stage 1 on a slab rank (generation of group j+1 overlaps the z pass of j)
    int p2p_ctas     = -1;  // ZPLT_P2P_CTAS: SMs of the z pass + exchange kernel (0 = as many as allowed, -1 = 64 — 56 on 2 ranks, 84 at N = 2048:
                            // best of the sweeps on 8 GPUs, profiles/r02_sweep_8gpu.jsonl); generation takes the rest
    int dit2048      = 1;   // ZPLT_DIT2048: 8-pencil decimation kernels for the N = 2048 z pass (122 -> 77 ms per rank of 8, local stores)
    int dit2048_emit = 0;   // ZPLT_DIT2048_EMIT: ... and for the N = 2048 y pass + emission (measured slower than the 4-pencil kernel: 41.9 vs 33.3 ms)
    int slab_ring    = 1;   // ZPLT_SLAB_RING: ring-prefetched forms of the slab-rank kernels
    int p2p_resident = 1;   // ZPLT_P2P_RESIDENT: one z pass + exchange launch for the whole of stage 1, gated by per-group flags
    int b2_layout    = -1;  // ZPLT_B2_LAYOUT: receive layout of the fused exchange: 0 = rows at their true y, 1 = per-source blocks,
                            // -1 = by measurement: per-source on >= 4 ranks up to N = 1024 (stage 1 at PPD=1024: 4 GPUs 26.0 -> 19.5 ms,
                            // 8 GPUs 14.3 -> 11.7 ms, NVLink 495/524 -> 662/642 GB/s per GPU; the y pass pays 0.7-1.4 ms for it), rows at
                            // their true y on 2 ranks and at N = 2048, where per-source blocks lose (8 GPUs: 195.7 vs 162.5 ms/step);
                            // profiles/r02_sweep_{4,8}gpu_receive_layout.jsonl
    int b2_pad       = 0;   // ZPLT_B2_PAD: complex elements added to the plane stride of layout 0 (<= ZPLT_B2_PAD_MAX)
    int p2p_helper   = 1;   // ZPLT_P2P_HELPER: a second launch of it on the SMs the generation kernels leave behind when they are done
    int gen_persist  = 1;   // ZPLT_GEN_PERSIST: persistent, software-pipelined generation kernel
};

// The z pass + exchange kernel of a slab rank can be resident for the whole of stage 1 and consume row groups as the generation
// kernels (another stream) complete them: after each generation kernel the host enqueues a stream-ordered memset of one flag.
struct GroupSync {
    const unsigned int *flags;  // [J] nonzero once the rows of group j have been generated; NULL: everything is ready
    unsigned int *err;          // set to 1 when a wait times out (the kernel then gives up instead of hanging the GPU)
    int J;                      // row groups covered by this launch
    unsigned int *counter;      // tile counter shared with another launch of the same pass (already reset by the caller), or NULL
};
// true when the z pass + exchange kernel for length N hands out its tiles through a device counter (ring-prefetched form), so that
// two launches can work through the same tiles together
bool fft_tiles_p2p_shares_tiles(int N, const Tuning &tn);

// Per-context launch resources: the work counters of the persistent kernels (a small rotating device array, so that
// launches in flight on different streams never share one) and the device's SM count.
struct LaunchRes {
    unsigned int *counters = nullptr;  // [64]
    int next_counter       = 0;
    int sms                = 0;
};

int fft_tile_T(int N);            // pencils per CTA used for length N (strided / row kernels)
int gen_xfft_T(int N, int na);    // pencils per CTA of the generation + x-FFT kernel
size_t fft_tile_smem(int N, int T);
// Fused mode generation + x-axis FFT, writes the whole [na][z][y][x] cube (skip_fft: the packed arrays as the
// kernel forms them, without the transform — introspection of the hot kernel).  cube = NULL: load the kernel, launch nothing.
int launch_gen_xfft(int N, int T, const GenParams &g, const SlabGeom &sg, cplx *cube, const cplx *tw, const Tuning &tn, LaunchRes &lr,
                    bool skip_fft, cudaStream_t st);
// In-place backward FFT of every pencil described by geom (tiles of T pencils).  Returns cudaError_t.
int launch_fft_tiles(int N, int T, cplx *data, const TileGeom &geom, const cplx *tw, const Tuning &tn, LaunchRes &lr, cudaStream_t st);
// z-axis FFT of a slab rank's stage-1 buffer with the exchange fused in: results are stored
// directly into every owner rank's stage-2 buffer (peer_recv[r], NVLink peer memory; NULL = discard).
int launch_fft_tiles_p2p(int N, int T, const cplx *b1, const SlabGeom &sg, cplx *const *peer_recv, const cplx *tw, const Tuning &tn,
                         LaunchRes &lr, const GroupSync &gs, cudaStream_t st);
// The same two launchers behind Tuning::dit2048: at N = 2048 they use the 8-pencil decimation kernels of
// zplt_fft2048_kernels.cu; otherwise they forward to the launchers above.
int launch_fft_tiles_any(int N, int T, cplx *data, const TileGeom &geom, const cplx *tw, const Tuning &tn, LaunchRes &lr, cudaStream_t st);
int launch_fft_tiles_p2p_any(int N, int T, const cplx *b1, const SlabGeom &sg, cplx *const *peer_recv, const cplx *tw,
                             const Tuning &tn, LaunchRes &lr, const GroupSync &gs, cudaStream_t st);
// N = 2048 y pass + emission with 8-pencil tiles (zplt_fft2048_kernels.cu); -1 = not applicable to this launch
int launch_fft2048_emit(const cplx *planes, long long z_first, long long nz, const EmitParams &ep, const cplx *tw, const Tuning &tn,
                        LaunchRes &lr, cudaStream_t st);
// y-axis FFT fused with record emission (cube: x and z already transformed; not modified).
int launch_fft_emit_strided(int N, int T, const cplx *cube, const SlabGeom &sg, long long z_first, long long nz,
                            const EmitParams &ep, const cplx *tw, const Tuning &tn, LaunchRes &lr, cudaStream_t st, int *launches);

// ppd not a power of two (zplt_generic_kernels.cu): one axis of the cube through Bluestein's algorithm on the length-M kernels
// (W = work array [M][Qb], w = chirp exp(i pi n^2/N), Bhat = transform of the conjugate chirp, twM = W_M table), unfused emission
int launch_bluestein_axis(cplx *cube, int N, int M, int na, int axis, cplx *W, long long Qb, const cplx *w, const cplx *Bhat,
                          const cplx *twM, const Tuning &tn, LaunchRes &lr, cudaStream_t st);
int launch_emit_plain(const cplx *cube, int N, long long z_first, long long nz, const EmitParams &ep, cudaStream_t st);

int launch_power_table(double *ptab, long long count, double fundamental2, int is_powerlaw, double index, int n,
                       const double *x, const double *y, const double *y2, double normalization, double smooth2,
                       cudaStream_t st);
int launch_generate(const GenParams &g, cplx *cube, cudaStream_t st);
// ZD_f_NL: M(k) table, phi_g(k) = D/M on the full lattice, the local transformation in configuration space
int launch_mfactor_table(double *mtab, const double *ptab, long long count, double fundamental2, double primordial_norm, double n_s,
                         double z_initial, double Omega_M, cudaStream_t st);
int launch_generate_phi(const GenParams &g, const SlabGeom &sg, cplx *phi, cudaStream_t st);
int launch_fnl_local(cplx *phi, int N, long long count, double f_NL, cudaStream_t st);
// slab ranks: rows y < N/2 of this rank's transformed potential planes [zl][y][x] to their owners' [z][slot][x] buffers
int launch_phi_return(const cplx *p2, const SlabGeom &sg, cplx *const *peer_p1, cudaStream_t st);
int launch_pcg_draws(const u128 *ystate0, const Affine *jump, long long n, uint64_t *out, cudaStream_t st);
int launch_mode_draws(const GenParams &g, long long n, const int *k, uint64_t *raw, double *u, cudaStream_t st);

// host helpers (zplt_tables.cpp part of api)
u128 pcg_seed_state(uint64_t seed);
Affine pcg_jump(unsigned __int128 delta);

}  // namespace zplt
