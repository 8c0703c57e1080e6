// C-ABI host side of pyxrd_b200 (see include/pyxrd_b200.h for the contract).
// Owns device memory for tables / parameters / results, builds the job lists and
// launches the kernels of kernels.cuh on one CUDA stream.  No CPU compute path.
#include "../../include/pyxrd_b200.h"
#include <cstdio>
#include "kernels.cuh"

#include <algorithm>
#include <cstdarg>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace {

struct DevBuf {
    void *ptr = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(ptr); }
};

struct PinBuf {
    void *ptr = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 4096;
        cudaError_t e = cudaMallocHost(&ptr, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        cap = 0;
    }
};

struct RankClass {
    int rank = 0, G = 0;
    int structure = 0;         // 1: every P of the class has the Reichweite >= 2 transition structure, 2: equal rows
                               // (Reichweite 0): k_phase_small<.., STRUCT>
    int max_x = 0;
    int max_ntop_cap = 0;
    size_t first = 0;          // position of this class in the device job-id list
    std::vector<int> jobs;     // ascending job ids (general batches)
    // populations expanded on the device: every candidate has the template's per_cand jobs of this class, the
    // device list holds them candidate by candidate; `jobs` is empty then
    size_t per_cand = 0;
    bool uniform = false;
    size_t count(int K) const { return uniform ? per_cand * (size_t)K : jobs.size(); }
    // [a, b) inside this class's device list of the jobs of candidates [k0, k1)
    void range(int k0, int k1, long long j0, long long j1, size_t &a, size_t &b) const {
        if (uniform) { a = per_cand * (size_t)k0; b = per_cand * (size_t)k1; return; }
        a = std::lower_bound(jobs.begin(), jobs.end(), (int)j0) - jobs.begin();
        b = std::lower_bound(jobs.begin(), jobs.end(), (int)j1) - jobs.begin();
    }
};

}  // namespace

// rank of the matrices a probability model produces for G components; -1 when the reference has no such model
static int pyxrd_prob_rank(int model, int G) {
    switch (model) {
        case PYXRD_PROB_R0: return (G >= 1 && G <= 6) ? G : -1;
        case PYXRD_PROB_R1G2: return G == 2 ? 2 : -1;
        case PYXRD_PROB_R2G2: return G == 2 ? 4 : -1;
        case PYXRD_PROB_R3G2: return G == 2 ? 8 : -1;
        case PYXRD_PROB_R1G3: return G == 3 ? 3 : -1;
        case PYXRD_PROB_R1G4: return G == 4 ? 4 : -1;
        case PYXRD_PROB_R2G3: return G == 3 ? 9 : -1;
        default: return -1;
    }
}

struct pyxrd_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;      // device -> host copies of finished candidate chunks (pyxrd_evaluate_host)
    cudaStream_t class_stream[3] = {};       // the job classes after the first run beside it (their tails overlap)
    cudaEvent_t fork_ev = nullptr, join_ev[3] = {};
    bool overlap_classes = true;             // pyxrd_set_option("overlap_classes", 0): one class after the other
    cudaEvent_t chunk_ev[16] = {};
    std::string error;
    int64_t launches = 0;
    int smem_optin = 0;
    int sm_count = 148;
    // measurement / experiment options, set explicitly through pyxrd_set_option (nothing is read from the environment)
    bool use_structure = true;   // "use_structure" 0: ignore the transition structure of P
    bool series_mma = true;      // "series_mma" 0: CUDA-core series kernels for the ranks that have a DMMA kernel
    bool use_layer_tables = true;   // "layer_tables" 0: every layer stack is evaluated inside the phase kernels
    int opt_tpb = 0;             // "optimizer_threads" 128 / 256: block size of k_mixture_optimize (0: by population size)
    int host_chunks = 0;         // "host_chunks" n: candidate chunks of pyxrd_evaluate_host (0: by population size)
    double opt_delta0 = 1e-2, opt_shrink = 0.5, opt_delta_min = 1e-7;   // weight-floor schedule of the optimiser
    int opt_single_from = 0;     // "optimizer_single_from": majoriser from which ONE Newton iteration is taken (0: never)
    // "optimizer_newton_tol": a Newton iteration that lowers the majoriser by less than this (relative) is the last of
    // its pass.  1e-5 keeps the excess over a long run on c4 AND on the ill-conditioned c5 (four similar phases) where
    // 1e-13 had it, 14 % / 9 % faster; 1e-4 and fixed single iterations after the first passes lose c5 (tools/opt_switch.py)
    double opt_newton_tol = 1e-5;
    int opt_coarse_until = 0, opt_coarse_stride = 1;   // "optimizer_coarse_until" / "optimizer_coarse_stride" (off)
    int opt_lane_solver = 4;     // "optimizer_lane_solver": dampings of the single-lane solver for 5 < D <= 12 (0: warp teams)

    // static
    bool have_static = false;
    StaticDev st{};
    std::vector<int> h_npts;
    std::vector<long long> h_off;
    std::vector<int> h_nwl;
    DevBuf d_npts, d_off, d_gonio, d_derived, d_theta, d_sin, d_invsin, d_stl, d_c2t2, d_corr, d_asf, d_nwl, d_wlfirst,
        d_wloff, d_wlfrac, d_wlidx, d_wlhi, d_wllo, d_obs, d_sel, d_atypes, d_stlov;

    // batch
    bool have_batch = false;
    bool have_intensities = false;
    BatchDev bt{};
    int NJ = 0;
    std::vector<RankClass> classes;
    DevBuf d_blob;           // all uploaded batch arrays, one allocation
    PinBuf h_blob;
    DevBuf d_tmpl;           // template candidate + solution vectors + binding table of a device-expanded population
    PinBuf h_tmpl;
    const int *jobids_dev = nullptr;   // inside d_blob
    DevBuf d_ph_nu, d_ph_uz, d_ph_uend, d_ph_useg, d_partial, d_phase_img;
    DevBuf d_comp_ilp, d_layer_tab;
    DevBuf d_atom_z, d_atom_plane, d_comp_np, d_phase_x, d_phase_valid, d_coef, d_int, d_int2, d_scaled, d_tot, d_res;
    int residual_method = PYXRD_RESIDUAL_RP;   // settings.RESIDUAL_METHOD (mixture.py:22-27,89)
    DevBuf d_selpos, d_selcount;
    std::vector<int> h_selcount;
    int wl_passes = 0;       // longest wavelength list of the specimens = number of ping-pong passes
    int wl_halo_l = 0, wl_halo_r = 0;   // how far below / above a point the passes of the distribution reach in total (k_patterns_residual)
    bool fuse_patterns = true;          // "fuse_patterns": pyxrd_compute_all runs all passes + residual in one kernel where the reach allows
    int jobs_per_cand = 0;
    PinBuf h_res;
    DevBuf d_leaf_a, d_leaf_b, d_leaf_c, d_leaf_d, d_sol, d_optres;
    bool have_opt = false;
    bool have_totals = false;
    int64_t int_count = 0, tot_count = 0;
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool ev_set[4] = {false, false, false, false};

    // what pyxrd_update_population needs of the resident device-expanded population
    struct Population {
        bool valid = false;
        ExpandDev e{};
        int ncap0 = 2, acap = 1;                     // coefficient stride without the automatic maxima / atom stride
        size_t x_off = 0;                            // offset of the solution vectors inside the staging blob
        std::vector<int32_t> t_phase_i;              // template parameter blocks (CSDS range, flags)
        std::vector<double> t_phase_d;
        std::vector<int> t_job_pb;
        std::vector<std::vector<int>> cls_jobs;      // template job ids of every class
        std::vector<int> cls_cap0;                   // class coefficient caps of the template itself
        std::vector<pyxrd_binding> bind;
        std::vector<uint8_t> auto_max;
        double factor = 2.5;
        long long NJ = 0;
        size_t njobids = 0;
        int M = 0;
        int n_layer_classes = 0;
        bool model_driven = false;
    } pop;
};

namespace {

int fail(pyxrd_ctx *ctx, const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->error = buf;
    return 1;
}

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e__ = (call);                                                                     \
        if (e__ != cudaSuccess)                                                                       \
            return fail(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#define LAUNCH_CHECK(name)                                                                            \
    do {                                                                                              \
        ctx->launches++;                                                                              \
        cudaError_t e__ = cudaGetLastError();                                                         \
        if (e__ != cudaSuccess) return fail(ctx, "launch of %s failed: %s", name, cudaGetErrorString(e__)); \
    } while (0)

template <class T>
int upload(pyxrd_ctx *ctx, DevBuf &buf, const T *src, size_t n) {
    CK(buf.ensure(std::max<size_t>(n, 1) * sizeof(T)));
    if (n) CK(cudaMemcpyAsync(buf.ptr, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// PhaseSmem, P, WG, coefficients, then (16-byte aligned) the sine / cosine table of px_sincos_tab_n
size_t small_sct_offset(int R, int G, int ntop_cap) {
    return align_up(sizeof(PhaseSmem) + sizeof(double) * (size_t)(R * R + G * R + ntop_cap + 2), 16);
}
size_t small_smem(int R, int G, int ntop_cap, bool table) { return small_sct_offset(R, G, ntop_cap) + (table ? sizeof(double2) * PX_SCT_N : 0); }
size_t staged_smem(int R, int G, int ntop_cap) {
    const int RB = (R + 3) / 4;
    return sizeof(PhaseSmem) + sizeof(double) * (size_t)(RB * R * 4 + G * R + (G * R & 1)) +
           sizeof(double2) * (size_t)R * PX_TPB + sizeof(double) * (size_t)(ntop_cap + 2);
}
size_t mma_smem(int R, int G, int ntop_cap) {
    return sizeof(PhaseSmem) + sizeof(double) * (size_t)(R * R + G * R + ((R * R + G * R) & 1)) +
           sizeof(double2) * (size_t)2 * G * PX_TPB + sizeof(double) * (size_t)(ntop_cap + 2);
}
size_t generic_smem(int R, int G, int ntop_cap) {
    return sizeof(PhaseSmem) + sizeof(double) * (size_t)(R * R + G * R + 1) +
           sizeof(double2) * (size_t)(2 * R + 16) * PX_GEN_TPB + sizeof(double) * (size_t)(ntop_cap + 2);
}

template <class K>
int set_smem(pyxrd_ctx *ctx, K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        if ((int)bytes > ctx->smem_optin) return fail(ctx, "kernel needs %zu B of shared memory, device offers %d", bytes, ctx->smem_optin);
        CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    }
    return 0;
}

#ifndef PX_SMALL_TILES
#define PX_SMALL_TILES 4     // tiles per block of the one-warp kernels (ranks 1, 2)
#endif
#ifndef PX_S22_PPT
#define PX_S22_PPT 4         // points per thread / blocks per SM of the rank-2, two-component kernel (c2)
#define PX_S22_MINB 16
#endif
#ifndef PX_WIDE_TILES
#define PX_WIDE_TILES 1      // ... of the kernels with 64 / 128 threads per block
#endif
template <int R, int G, int PPT, int TPB, int MINB = 0, int STRUCT = 0, int TILES = 1>
int launch_small(pyxrd_ctx *ctx, const RankClass &rc, const int *jobs_dev, size_t njobs, const BatchDev &bt) {
    static_assert((PPT - 1) * TPB <= PX_TABLE_PAD, "table padding too small");
    const size_t smem = small_smem(R, G, rc.max_ntop_cap, PX_SCT_SMEM && MINB > 0);
    if (set_smem(ctx, k_phase_small<R, G, PPT, TPB, MINB, STRUCT, TILES>, smem)) return 1;
    dim3 grid((unsigned)njobs, (unsigned)((rc.max_x + TPB * PPT * TILES - 1) / (TPB * PPT * TILES)));
    BatchDev b2 = bt;
    b2.img_ncopy = std::min(bt.ncap, rc.max_ntop_cap + 2);      // what the launch's shared memory holds
    b2.img_sct_off = (int)small_sct_offset(R, G, rc.max_ntop_cap);
    k_phase_small<R, G, PPT, TPB, MINB, STRUCT, TILES><<<grid, TPB, smem, ctx->stream>>>(ctx->st, b2, jobs_dev);
    LAUNCH_CHECK("k_phase_small");
    return 0;
}

template <int R, int G>
int launch_staged(pyxrd_ctx *ctx, const RankClass &rc, const int *jobs_dev, size_t njobs, const BatchDev &bt) {
    const size_t smem = staged_smem(R, G, rc.max_ntop_cap);
    if (set_smem(ctx, k_phase_staged<R, G>, smem)) return 1;
    dim3 grid((unsigned)njobs, (unsigned)((rc.max_x + PX_TPB - 1) / PX_TPB));
    k_phase_staged<R, G><<<grid, PX_TPB, smem, ctx->stream>>>(ctx->st, bt, jobs_dev);
    LAUNCH_CHECK("k_phase_staged");
    return 0;
}

template <int R, int G, int MTC>
int launch_mma(pyxrd_ctx *ctx, const RankClass &rc, const int *jobs_dev, size_t njobs, const BatchDev &bt) {
    const size_t smem = mma_smem(R, G, rc.max_ntop_cap);
    if (set_smem(ctx, k_phase_mma<R, G, MTC>, smem)) return 1;
    dim3 grid((unsigned)njobs, (unsigned)((rc.max_x + PX_TPB - 1) / PX_TPB));
    k_phase_mma<R, G, MTC><<<grid, PX_TPB, smem, ctx->stream>>>(ctx->st, bt, jobs_dev);
    LAUNCH_CHECK("k_phase_mma");
    return 0;
}

int launch_generic(pyxrd_ctx *ctx, const RankClass &rc, const int *jobs_dev, size_t njobs, const BatchDev &bt) {
    const size_t smem = generic_smem(rc.rank, rc.G, rc.max_ntop_cap);
    if (set_smem(ctx, k_phase_generic, smem)) return 1;
    dim3 grid((unsigned)njobs, (unsigned)((rc.max_x + PX_GEN_TPB - 1) / PX_GEN_TPB));
    k_phase_generic<<<grid, PX_GEN_TPB, smem, ctx->stream>>>(ctx->st, bt, jobs_dev);
    LAUNCH_CHECK("k_phase_generic");
    return 0;
}

int launch_raw(pyxrd_ctx *ctx, const RankClass &rc, const int *jobs_dev, size_t njobs, const BatchDev &bt) {
    dim3 grid((unsigned)njobs, (unsigned)((rc.max_x + 127) / 128));
    k_phase_raw<<<grid, 128, 0, ctx->stream>>>(ctx->st, bt, jobs_dev);
    LAUNCH_CHECK("k_phase_raw");
    return 0;
}

int launch_class(pyxrd_ctx *ctx, const RankClass &rc, const int *jobs_dev, size_t njobs, const BatchDev &bt) {
    if (rc.rank == 0 && rc.G == 0) return launch_raw(ctx, rc, jobs_dev, njobs, bt);
#define SMALL(R_, G_, PPT_, TPB_) if (rc.rank == R_ && rc.G == G_) return launch_small<R_, G_, PPT_, TPB_, 0, 0, PX_WIDE_TILES>(ctx, rc, jobs_dev, njobs, bt)
    // one-warp kernels: four tiles per block where the launch is many waves deep (staging paid once per 512 points), two
    // where it is not (short patterns of a small population: the last blocks of a pattern are nearly empty and the tail of
    // the launch counts; c4 phase stage 0.325 -> 0.312 ms)
    const bool deep = (double)njobs * ((rc.max_x + 128 * PX_SMALL_TILES - 1) / (128 * PX_SMALL_TILES)) >= 6.0 * 16.0 * ctx->sm_count;
#define SMALLB(R_, G_, PPT_, TPB_, MINB_) if (rc.rank == R_ && rc.G == G_) return deep ? launch_small<R_, G_, PPT_, TPB_, MINB_, 0, PX_SMALL_TILES>(ctx, rc, jobs_dev, njobs, bt) \
                                                                                    : launch_small<R_, G_, PPT_, TPB_, MINB_, 0, PX_SMALL_TILES / 2>(ctx, rc, jobs_dev, njobs, bt)
#define STAGED(R_, G_) if (rc.rank == R_ && rc.G == G_) return launch_staged<R_, G_>(ctx, rc, jobs_dev, njobs, bt)
    if (rc.structure == 2 && ctx->use_structure) {
        // Reichweite 0 with G components: a scalar series whatever G is
        if (rc.rank == 2 && rc.G == 2) return deep ? launch_small<2, 2, 4, 32, 16, 2, PX_SMALL_TILES>(ctx, rc, jobs_dev, njobs, bt)
                                                   : launch_small<2, 2, 4, 32, 16, 2, PX_SMALL_TILES / 2>(ctx, rc, jobs_dev, njobs, bt);
        if (rc.rank == 3 && rc.G == 3) return launch_small<3, 3, 2, 128, 0, 2, PX_WIDE_TILES>(ctx, rc, jobs_dev, njobs, bt);
        if (rc.rank == 4 && rc.G == 4) return launch_small<4, 4, 2, 128, 0, 2, PX_WIDE_TILES>(ctx, rc, jobs_dev, njobs, bt);
        if (rc.rank == 5 && rc.G == 5) return launch_small<5, 5, 1, 128, 0, 2, PX_WIDE_TILES>(ctx, rc, jobs_dev, njobs, bt);
        if (rc.rank == 6 && rc.G == 6) return launch_small<6, 6, 1, 128, 0, 2, PX_WIDE_TILES>(ctx, rc, jobs_dev, njobs, bt);
    }
    if (rc.structure == 1 && ctx->use_structure) {
        // structured transition matrices (Reichweite >= 2 models): G instead of R multiply-adds per state
        // (shapes from sweeps on B200)
        if (rc.rank == 4 && rc.G == 2) return launch_small<4, 2, 2, 128, 0, 1, PX_WIDE_TILES>(ctx, rc, jobs_dev, njobs, bt);
        if (rc.rank == 8 && rc.G == 2) return launch_small<8, 2, 2, 64, 8, 1, PX_WIDE_TILES>(ctx, rc, jobs_dev, njobs, bt);   // 3.64 ms on c5; 1 x 128: 4.02
        if (rc.rank == 9 && rc.G == 3) return launch_small<9, 3, 1, 128, 0, 1, PX_WIDE_TILES>(ctx, rc, jobs_dev, njobs, bt);
        // larger ranks: the in-place series needs the state vector once (254 registers at rank 27); c3s 1.52 ms per 256
        // candidates, 4.07 ms with the vector staged through shared memory, 7.89 ms with the dense DMMA kernel
        if (rc.rank == 27 && rc.G == 3) return launch_small<27, 3, 1, 128, 0, 1, PX_WIDE_TILES>(ctx, rc, jobs_dev, njobs, bt);
        if (rc.rank == 16 && rc.G == 2) return launch_small<16, 2, 1, 128, 0, 1, PX_WIDE_TILES>(ctx, rc, jobs_dev, njobs, bt);
        if (rc.rank == 16 && rc.G == 4) return launch_small<16, 4, 1, 128, 0, 1, PX_WIDE_TILES>(ctx, rc, jobs_dev, njobs, bt);
    }
    // (points per thread, threads per block, min blocks per SM) from sweeps on B200 (profiles/r1_summary.md)
    SMALLB(1, 1, 4, 32, 0); SMALLB(2, 2, PX_S22_PPT, 32, PX_S22_MINB); SMALLB(2, 1, 4, 32, 16);
    SMALL(3, 3, 2, 128); SMALL(4, 4, 2, 128);  // R0 / R1
    SMALL(5, 5, 1, 128); SMALL(6, 6, 1, 128);
    SMALL(3, 1, 2, 128); SMALL(4, 1, 2, 128);
#define MMA(R_, G_, MTC_) if (rc.rank == R_ && rc.G == G_ && ctx->series_mma) return launch_mma<R_, G_, MTC_>(ctx, rc, jobs_dev, njobs, bt)
    MMA(8, 2, 4); MMA(16, 2, 2); MMA(16, 4, 2); MMA(27, 3, 1);                        // FP64 tensor cores
    SMALL(4, 2, 2, 128); SMALL(8, 2, 1, 128); SMALL(9, 3, 1, 128);                                            // R2G2, R3G2, R2G3
    STAGED(16, 2); STAGED(16, 4); STAGED(27, 3);
#undef SMALL
#undef SMALLB
#undef STAGED
#undef MMA
    return launch_generic(ctx, rc, jobs_dev, njobs, bt);
}

static int read_back(pyxrd_ctx *ctx, const void *dev, double *out, int64_t count, int64_t have, const char *what) {
    if (!dev || count > have || count < 0) return fail(ctx, "read %s: asked for %lld doubles, buffer holds %lld", what, (long long)count, (long long)have);
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(out, dev, sizeof(double) * (size_t)count, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}


template <int MB>
static int launch_optimize(pyxrd_ctx *ctx, const OptParams &op) {
    // one wave of 256-thread blocks (2 per SM) when the population fits it, else 128-thread blocks (4 per SM)
    int tpb = ctx->bt.K <= 2 * ctx->sm_count ? 256 : 128;
    if (ctx->opt_tpb) tpb = ctx->opt_tpb;
    if (tpb == 256) {
        k_mixture_optimize<MB, 256><<<ctx->bt.K, 256, 0, ctx->stream>>>(ctx->st, ctx->bt, op, ctx->d_sol.as<double>(),
                                                                           ctx->d_optres.as<double>());
    } else {
        k_mixture_optimize<MB, 128><<<ctx->bt.K, 128, 0, ctx->stream>>>(ctx->st, ctx->bt, op, ctx->d_sol.as<double>(),
                                                                           ctx->d_optres.as<double>());
    }
    LAUNCH_CHECK("k_mixture_optimize");
    return 0;
}

}  // namespace

extern "C" {

int pyxrd_create(int device, pyxrd_ctx **out) {
    if (!out) return 1;
    *out = nullptr;
    pyxrd_ctx *ctx = new pyxrd_ctx();
    ctx->device = device;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0 || device >= count) {
        fprintf(stderr, "pyxrd_b200: no usable CUDA device (%s); there is no CPU fallback\n",
                e == cudaSuccess ? "device index out of range" : cudaGetErrorString(e));
        delete ctx;
        return 2;
    }
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return 3;
    }
    ctx->stream = ctx->own_stream;
    cudaDeviceGetAttribute(&ctx->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    for (int i = 0; i < 8; ++i) cudaEventCreate(&ctx->ev[i]);
    cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 3; ++i) {
        cudaStreamCreateWithFlags(&ctx->class_stream[i], cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&ctx->join_ev[i], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming);
    {   // np.blackman(31) / np.blackman(31).sum()  (math_tools.py:93-98 with half_window_len = 15)
        double w[31], sum = 0.0;
        for (int i = 0; i < 31; ++i) {
            const double n = 2.0 * i - 30.0;      // arange(1 - M, M, 2)
            w[i] = 0.42 + 0.5 * cos(PX_PI * n / 30.0) + 0.08 * cos(2.0 * PX_PI * n / 30.0);
            sum += w[i];
        }
        for (int i = 0; i < 31; ++i) w[i] /= sum;
        cudaMemcpyToSymbol(PX_BLACKMAN, w, sizeof(w));
    }
    for (int i = 0; i < 16; ++i) cudaEventCreateWithFlags(&ctx->chunk_ev[i], cudaEventDisableTiming);
    *out = ctx;
    return 0;
}

void pyxrd_destroy(pyxrd_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->d_npts, &ctx->d_off, &ctx->d_gonio, &ctx->d_derived, &ctx->d_theta, &ctx->d_sin, &ctx->d_invsin, &ctx->d_stl,
                      &ctx->d_c2t2, &ctx->d_corr, &ctx->d_asf, &ctx->d_nwl, &ctx->d_wlfirst, &ctx->d_wloff, &ctx->d_wlfrac,
                      &ctx->d_wlidx, &ctx->d_wlhi, &ctx->d_wllo, &ctx->d_obs, &ctx->d_sel, &ctx->d_atypes, &ctx->d_stlov, &ctx->d_blob, &ctx->d_tmpl,
                      &ctx->d_partial, &ctx->d_phase_img, &ctx->d_ph_nu, &ctx->d_ph_uz, &ctx->d_ph_uend, &ctx->d_ph_useg, &ctx->d_atom_z, &ctx->d_atom_plane, &ctx->d_comp_np, &ctx->d_phase_x, &ctx->d_phase_valid, &ctx->d_coef, &ctx->d_int, &ctx->d_scaled, &ctx->d_tot,
                      &ctx->d_res, &ctx->d_comp_ilp, &ctx->d_layer_tab, &ctx->d_int2, &ctx->d_selpos, &ctx->d_selcount, &ctx->d_leaf_a, &ctx->d_leaf_b, &ctx->d_leaf_c, &ctx->d_leaf_d, &ctx->d_sol, &ctx->d_optres};
    for (DevBuf *b : bufs) b->release();
    ctx->h_blob.release();
    ctx->h_tmpl.release();
    ctx->h_res.release();
    for (int i = 0; i < 8; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 16; ++i) if (ctx->chunk_ev[i]) cudaEventDestroy(ctx->chunk_ev[i]);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (int i = 0; i < 3; ++i) {
        if (ctx->class_stream[i]) cudaStreamDestroy(ctx->class_stream[i]);
        if (ctx->join_ev[i]) cudaEventDestroy(ctx->join_ev[i]);
    }
    if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char *pyxrd_last_error(const pyxrd_ctx *ctx) { return ctx ? ctx->error.c_str() : "null context"; }

int pyxrd_set_stream(pyxrd_ctx *ctx, void *cuda_stream) {
    if (!ctx) return 1;
    ctx->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
    return 0;
}

int pyxrd_synchronize(pyxrd_ctx *ctx) {
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int64_t pyxrd_launch_count(const pyxrd_ctx *ctx) { return ctx ? ctx->launches : 0; }

int pyxrd_set_option(pyxrd_ctx *ctx, const char *name, double value) {
    if (!ctx || !name) return 1;
    const std::string key(name);
    if (key == "use_structure") ctx->use_structure = value != 0.0;
    else if (key == "series_mma") ctx->series_mma = value != 0.0;
    else if (key == "overlap_classes") ctx->overlap_classes = value != 0.0;
    else if (key == "fuse_patterns") ctx->fuse_patterns = value != 0.0;
    else if (key == "layer_tables") ctx->use_layer_tables = value != 0.0;
    else if (key == "optimizer_threads") {
        if (value != 0.0 && value != 128.0 && value != 256.0) return fail(ctx, "set_option: optimizer_threads is 0, 128 or 256");
        ctx->opt_tpb = (int)value;
    } else if (key == "optimizer_coarse_until") {
        ctx->opt_coarse_until = value > 0.0 ? (int)value : 0;
    } else if (key == "optimizer_coarse_stride") {
        ctx->opt_coarse_stride = value >= 1.0 ? (int)value : 1;
    } else if (key == "optimizer_newton_tol") {
        ctx->opt_newton_tol = value;
    } else if (key == "optimizer_lane_solver") {
        if (value != 0.0 && value != 4.0 && value != 8.0) return fail(ctx, "set_option: optimizer_lane_solver is 0, 4 or 8");
        ctx->opt_lane_solver = (int)value;
    } else if (key == "optimizer_single_from") {
        ctx->opt_single_from = value > 0.0 ? (int)value : 0;
    } else if (key == "host_chunks") {
        if (value < 0.0 || value > 16.0) return fail(ctx, "set_option: host_chunks is 0 .. 16");
        ctx->host_chunks = (int)value;
    } else if (key == "optimizer_delta0" || key == "optimizer_shrink" || key == "optimizer_delta_min") {
        if (!(value > 0.0) || (key == "optimizer_shrink" && value > 1.0)) return fail(ctx, "set_option: %s must be positive (shrink <= 1)", name);
        (key == "optimizer_delta0" ? ctx->opt_delta0 : key == "optimizer_shrink" ? ctx->opt_shrink : ctx->opt_delta_min) = value;
    } else {
        return fail(ctx, "set_option: unknown option '%s'", name);
    }
    return 0;
}

int pyxrd_set_static(pyxrd_ctx *ctx, const pyxrd_static_desc *d) {
    if (!ctx || !d) return 1;
    CK(cudaSetDevice(ctx->device));
    const int S = d->n_specimens, T = d->n_atom_types, Z = d->n_patterns;
    if (S <= 0 || Z <= 0 || T < 0) return fail(ctx, "set_static: need S > 0 and Z > 0 (S=%d Z=%d T=%d)", S, Z, T);
    ctx->have_static = false;
    ctx->have_batch = false;
    ctx->pop.valid = false;
    ctx->h_npts.assign(d->spec_npts, d->spec_npts + S);
    ctx->h_nwl.assign(d->spec_nwl, d->spec_nwl + S);
    ctx->h_off.assign(S + 1, 0);
    int maxX = 0;
    std::vector<int> wlfirst(S);
    std::vector<long long> wloff(S);
    long long nwl_total = 0, wl_pts = 0;
    for (int s = 0; s < S; ++s) {
        if (d->spec_npts[s] < 0) return fail(ctx, "set_static: negative point count");
        ctx->h_off[s + 1] = ctx->h_off[s] + d->spec_npts[s];
        maxX = std::max(maxX, d->spec_npts[s]);
        wlfirst[s] = (int)nwl_total;
        wloff[s] = wl_pts;
        nwl_total += d->spec_nwl[s];
        wl_pts += (long long)d->spec_nwl[s] * d->spec_npts[s];
    }
    const long long sumX = ctx->h_off[S];
    {
        // reach of the gathers of every pass of the wavelength distribution (k_patterns_residual): pass j of specimen s
        // reads at most r-_j points below and r+_j above the point it writes; the passes add up
        ctx->wl_halo_l = ctx->wl_halo_r = 0;
        for (int s = 0; s < S; ++s) {
            const int X = d->spec_npts[s];
            long long below = 0, above = 0;
            for (int j = 0; j < d->spec_nwl[s]; ++j) {
                const int32_t *idx = d->wl_index + wloff[s] + (long long)j * X;
                int lo = 0, hi = 0;
                for (int x = 0; x < X; ++x)
                    if (idx[x] >= 0) { lo = std::max(lo, x - (idx[x] - 1)); hi = std::max(hi, idx[x] - x); }
                below += lo; above += hi;
            }
            ctx->wl_halo_l = (int)std::max<long long>(ctx->wl_halo_l, std::min<long long>(below, 1 << 20));
            ctx->wl_halo_r = (int)std::max<long long>(ctx->wl_halo_r, std::min<long long>(above, 1 << 20));
        }
    }
    if (upload(ctx, ctx->d_npts, d->spec_npts, S)) return 1;
    if (upload(ctx, ctx->d_off, ctx->h_off.data(), S + 1)) return 1;
    if (upload(ctx, ctx->d_gonio, d->spec_gonio, (size_t)S * PYXRD_GONIO_STRIDE)) return 1;
    if (upload(ctx, ctx->d_theta, d->theta, sumX)) return 1;
    if (upload(ctx, ctx->d_nwl, d->spec_nwl, S)) return 1;
    if (upload(ctx, ctx->d_wlfirst, wlfirst.data(), S)) return 1;
    if (upload(ctx, ctx->d_wloff, wloff.data(), S)) return 1;
    if (upload(ctx, ctx->d_wlfrac, d->wl_fraction, nwl_total)) return 1;
    if (upload(ctx, ctx->d_wlidx, d->wl_index, wl_pts)) return 1;
    if (upload(ctx, ctx->d_wlhi, d->wl_w_hi, wl_pts)) return 1;
    if (upload(ctx, ctx->d_wllo, d->wl_w_lo, wl_pts)) return 1;
    if (upload(ctx, ctx->d_obs, d->observed, (size_t)Z * sumX)) return 1;
    if (upload(ctx, ctx->d_sel, d->selected, sumX)) return 1;
    if (upload(ctx, ctx->d_atypes, d->atom_types, (size_t)T * PYXRD_ATYPE_STRIDE)) return 1;
    if (d->stl_override && upload(ctx, ctx->d_stlov, d->stl_override, sumX)) return 1;
    // the upload() calls above read caller memory asynchronously only for pinned sources; for
    // pageable sources cudaMemcpyAsync returns after staging, so the caller may free them now
    CK(ctx->d_derived.ensure(sizeof(double) * 8 * S));
    CK(cudaMemsetAsync(ctx->d_derived.ptr, 0, sizeof(double) * 8 * S, ctx->stream));
    // every per-point table carries PX_TABLE_PAD zeroed doubles of slack: threads of k_phase_small
    // that own several points load (and discard) values past the end of the last pattern
    {
        DevBuf *tabs[] = {&ctx->d_sin, &ctx->d_stl, &ctx->d_c2t2, &ctx->d_corr, &ctx->d_asf, &ctx->d_invsin};
        const long long counts[] = {sumX, sumX, sumX, sumX, sumX * std::max(T, 1), sumX};
        for (int i = 0; i < 6; ++i) {
            const size_t bytes = sizeof(double) * (size_t)(counts[i] + PX_TABLE_PAD);
            CK(tabs[i]->ensure(bytes));
            CK(cudaMemsetAsync(tabs[i]->ptr, 0, bytes, ctx->stream));
        }
    }

    StaticDev &st = ctx->st;
    st.S = S; st.T = T; st.Z = Z; st.maxX = maxX; st.sumX = sumX;
    st.spec_npts = ctx->d_npts.as<int>();
    st.spec_off = ctx->d_off.as<long long>();
    st.gonio = ctx->d_gonio.as<double>();
    st.derived = ctx->d_derived.as<double>();
    st.theta = ctx->d_theta.as<double>();
    st.sin_t = ctx->d_sin.as<double>();
    st.inv_sin = ctx->d_invsin.as<double>();
    st.stl = ctx->d_stl.as<double>();
    st.c2t2 = ctx->d_c2t2.as<double>();
    st.corr = ctx->d_corr.as<double>();
    st.asf = ctx->d_asf.as<double>();
    st.spec_nwl = ctx->d_nwl.as<int>();
    st.wl_first = ctx->d_wlfirst.as<int>();
    st.wl_off = ctx->d_wloff.as<long long>();
    st.wl_fraction = ctx->d_wlfrac.as<double>();
    st.wl_index = ctx->d_wlidx.as<int>();
    st.wl_w_hi = ctx->d_wlhi.as<double>();
    st.wl_w_lo = ctx->d_wllo.as<double>();
    st.observed = ctx->d_obs.as<double>();
    st.selected = ctx->d_sel.as<unsigned char>();
    st.atypes = ctx->d_atypes.as<double>();

    k_specimen_derived<<<(S + 63) / 64, 64, 0, ctx->stream>>>(S, st.gonio, ctx->d_derived.as<double>());
    LAUNCH_CHECK("k_specimen_derived");
    if (maxX > 0) {
        dim3 grid((maxX + 255) / 256, S);
        k_specimen_tables<<<grid, 256, 0, ctx->stream>>>(st, ctx->d_sin.as<double>(), ctx->d_invsin.as<double>(), ctx->d_stl.as<double>(),
                                                         ctx->d_c2t2.as<double>(), ctx->d_corr.as<double>(),
                                                         d->stl_override ? ctx->d_stlov.as<double>() : nullptr);
        LAUNCH_CHECK("k_specimen_tables");
        if (T > 0) {
            dim3 g2((unsigned)((sumX + 255) / 256), T);
            k_asf_table<<<g2, 256, 0, ctx->stream>>>(st, ctx->d_asf.as<double>());
            LAUNCH_CHECK("k_asf_table");
        }
    }
    k_observed_norm<<<S, 256, 0, ctx->stream>>>(st, ctx->d_derived.as<double>());
    LAUNCH_CHECK("k_observed_norm");
    CK(ctx->d_selpos.ensure(sizeof(int) * std::max<long long>(sumX, 1)));
    CK(ctx->d_selcount.ensure(sizeof(int) * S));
    k_selected_positions<<<S, 32, 0, ctx->stream>>>(st, ctx->d_selpos.as<int>(), ctx->d_selcount.as<int>());
    LAUNCH_CHECK("k_selected_positions");
    ctx->h_selcount.assign(S, 0);
    CK(cudaMemcpyAsync(ctx->h_selcount.data(), ctx->d_selcount.ptr, sizeof(int) * S, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_static = true;
    return 0;
}

}  // extern "C"

namespace {

// parts of the batch blob (one device allocation, identical layout for uploaded and device-expanded batches)
enum { P_PHASE_I, P_PHASE_D, P_PHASE_COMP, P_MATRICES, P_COMP_D, P_COMP_I, P_ATOM_D, P_ATOM_T, P_JOB_PB, P_JOB_SPEC,
       P_JOB_OUT, P_JOBIDS, P_FRAC, P_SCALES, P_BG, P_PMODEL, P_PPARAM, P_COMP_CLASS, P_CLASS_REP, P_COUNT };

#define PX_MAX_LAYER_CLASSES 64
#define PX_MIN_CLASS_MEMBERS 3     // a layer stack is tabulated when at least this many components of the batch share it

struct BatchSizes {
    int K = 0, M = 0, NP = 0, NC = 0, NA = 0, NPC = 0;
    int64_t NM = 0;
    long long NJ = 0;
    size_t njobids = 0;
    bool model_driven = false;
    int n_classes = 0;
};

struct BlobLayout {
    size_t off[P_COUNT], bytes[P_COUNT], total = 0;
};

BlobLayout plan_blob(const BatchSizes &z, int S) {
    BlobLayout L;
    const size_t bytes[P_COUNT] = {
        sizeof(int32_t) * (size_t)z.NP * PYXRD_PHASE_I_STRIDE, sizeof(double) * (size_t)z.NP * PYXRD_PHASE_D_STRIDE,
        sizeof(int32_t) * (size_t)z.NPC, sizeof(double) * (size_t)z.NM, sizeof(double) * (size_t)z.NC * PYXRD_COMP_D_STRIDE,
        sizeof(int32_t) * (size_t)z.NC * PYXRD_COMP_I_STRIDE, sizeof(double) * (size_t)z.NA * PYXRD_ATOM_D_STRIDE,
        sizeof(int32_t) * (size_t)z.NA, sizeof(int) * (size_t)z.NJ, sizeof(int) * (size_t)z.NJ, sizeof(long long) * (size_t)z.NJ,
        sizeof(int) * z.njobids, sizeof(double) * (size_t)z.K * z.M, sizeof(double) * (size_t)z.K * S, sizeof(double) * (size_t)z.K * S,
        z.model_driven ? sizeof(int32_t) * (size_t)z.NP : 0, z.model_driven ? sizeof(double) * (size_t)z.NP * PYXRD_PROB_STRIDE : 0,
        sizeof(int32_t) * (size_t)z.NC, sizeof(int32_t) * PX_MAX_LAYER_CLASSES};
    for (int i = 0; i < P_COUNT; ++i) {
        L.off[i] = L.total;
        L.bytes[i] = bytes[i];
        L.total = align_up(L.total + std::max<size_t>(bytes[i], 8), 256);
    }
    return L;
}

// parameter blocks of a batch descriptor are well formed; *ncap = coefficient stride (largest CSDS maximum + 2)
int validate_batch(pyxrd_ctx *ctx, const pyxrd_batch_desc *d, int *ncap_out, int *acap_out) {
    const StaticDev &st = ctx->st;
    const int NP = d->n_phases, NC = d->n_components, NA = d->n_atoms;
    int ncap = 2, acap = 1;
    for (int p = 0; p < NP; ++p) {
        const int32_t *pi = d->phase_i + (size_t)p * PYXRD_PHASE_I_STRIDE;
        if (!(pi[4] & PYXRD_FLAG_VALID_PROBS)) continue;
        const int G = pi[0], r = pi[1];
        if (pi[4] & PYXRD_FLAG_RAW_PATTERN) {           // RawPatternPhase: pool entry = raw_x[n] | raw_y[n]
            if (G != 0 || r != 0 || pi[2] < 0 || pi[6] < 0 || (int64_t)pi[6] + 2 * (int64_t)pi[2] > d->n_matrix)
                return fail(ctx, "set_batch: raw-pattern phase %d is malformed (n=%d, offset=%d, pool=%lld)", p, pi[2], pi[6], (long long)d->n_matrix);
            continue;
        }
        if (G < 1 || G > PYXRD_MAX_COMPONENTS || r < 1 || r > PYXRD_MAX_RANK || r % G != 0)
            return fail(ctx, "set_batch: phase %d has unsupported G=%d rank=%d (need 1<=G<=%d, rank<=%d, rank %% G == 0)",
                        p, G, r, PYXRD_MAX_COMPONENTS, PYXRD_MAX_RANK);
        if (pi[2] < 0 || pi[3] < pi[2]) return fail(ctx, "set_batch: phase %d CSDS range [%d, %d] unsupported", p, pi[2], pi[3]);
        if (pi[6] < 0 || (int64_t)pi[6] + 2 * (int64_t)r * r > d->n_matrix)
            return fail(ctx, "set_batch: phase %d: W / P at offset %d leave the matrix pool (%lld doubles)", p, pi[6], (long long)d->n_matrix);
        if (d->prob_model && d->prob_params && d->prob_model[p] != PYXRD_PROB_FIXED) {
            const int m = d->prob_model[p];
            if (pyxrd_prob_rank(m, G) != r)
                return fail(ctx, "set_batch: phase %d: probability model %d does not fit G=%d rank=%d", p, m, G, r);
        }
        ncap = std::max(ncap, pi[3] + 2);
        int atoms = 0;
        if (pi[5] < 0 || pi[5] + G > d->n_phase_comp) return fail(ctx, "set_batch: phase %d: component list leaves phase_comp", p);
        for (int c = 0; c < G; ++c) {
            const int ci = d->phase_comp[pi[5] + c];
            if (ci < 0 || ci >= NC) return fail(ctx, "set_batch: phase %d references component %d of %d", p, ci, NC);
            atoms += d->comp_i[(size_t)ci * PYXRD_COMP_I_STRIDE + 1] + d->comp_i[(size_t)ci * PYXRD_COMP_I_STRIDE + 2];
        }
        if (atoms > PX_MAX_ATOMS) return fail(ctx, "set_batch: phase %d has %d atoms (limit %d)", p, atoms, PX_MAX_ATOMS);
        acap = std::max(acap, atoms);
    }
    for (int a = 0; a < NA; ++a)
        if (d->atom_type[a] >= st.T) return fail(ctx, "set_batch: atom %d references atom type %d of %d", a, d->atom_type[a], st.T);
    *ncap_out = ncap;
    *acap_out = acap;
    return 0;
}

// Layer classes of the components of a descriptor: components whose layer rows (atom type, content, position of every
// layer atom, in packed order) are identical share a class.  `bound`: components a refinable drives (population
// templates) never join a class; `weight`: how many components of the final batch every listed component stands for.
// Classes with fewer than PX_MIN_CLASS_MEMBERS members (or beyond PX_MAX_LAYER_CLASSES) are dissolved (-1): their layers
// are evaluated inside the phase kernels.  -> comp_class[NC], class_rep[n_classes]
int layer_classes(const pyxrd_batch_desc *d, const std::vector<char> *bound, int weight, std::vector<int32_t> &comp_class,
                  std::vector<int32_t> &class_rep) {
    const int NC = d->n_components;
    comp_class.assign((size_t)std::max(NC, 1), -1);
    class_rep.clear();
    std::map<std::string, int> seen;
    std::vector<int> members;
    std::string key;
    for (int c = 0; c < NC; ++c) {
        if (bound && (*bound)[c]) continue;
        const int32_t *ci = d->comp_i + (size_t)c * PYXRD_COMP_I_STRIDE;
        const int a0 = ci[0], nl = ci[1];
        if (nl <= 0 || a0 < 0 || a0 + nl > d->n_atoms) continue;
        key.assign(reinterpret_cast<const char *>(d->atom_d + (size_t)a0 * PYXRD_ATOM_D_STRIDE), sizeof(double) * PYXRD_ATOM_D_STRIDE * nl);
        key.append(reinterpret_cast<const char *>(d->atom_type + a0), sizeof(int32_t) * nl);
        auto it = seen.find(key);
        if (it == seen.end()) {
            if ((int)class_rep.size() >= PX_MAX_LAYER_CLASSES) continue;
            it = seen.emplace(key, (int)class_rep.size()).first;
            class_rep.push_back(c);
            members.push_back(0);
        }
        comp_class[c] = it->second;
        members[it->second] += weight;
    }
    // dissolve small classes, renumber the rest
    std::vector<int> renumber(class_rep.size(), -1);
    std::vector<int32_t> kept;
    for (size_t i = 0; i < class_rep.size(); ++i)
        if (members[i] >= PX_MIN_CLASS_MEMBERS) { renumber[i] = (int)kept.size(); kept.push_back(class_rep[i]); }
    for (int c = 0; c < NC; ++c)
        if (comp_class[c] >= 0) comp_class[c] = renumber[comp_class[c]];
    class_rep.swap(kept);
    return (int)class_rep.size();
}

// job tables [K][S][Z][M] of a descriptor and the job lists per (rank, G) class, heavy classes first
int build_jobs(pyxrd_ctx *ctx, const pyxrd_batch_desc *d, std::vector<int> &job_pb, std::vector<int> &job_spec,
               std::vector<long long> &job_out, std::vector<int> &jobids) {
    const StaticDev &st = ctx->st;
    const int K = d->n_candidates, M = d->n_phase_slots, NP = d->n_phases, S = st.S, Z = st.Z;
    const long long cand_stride = (long long)M * Z * st.sumX;
    const long long NJ = (long long)K * S * Z * M;
    if (NJ > 0x7fffffffLL) return fail(ctx, "set_batch: too many patterns (%lld)", NJ);
    job_pb.assign((size_t)NJ, 0);
    job_spec.assign((size_t)NJ, 0);
    job_out.assign((size_t)NJ, 0);
    // phases whose P has the Reichweite >= 2 transition structure: known from the probability model, or checked
    // entry by entry (zero outside the columns (I mod r/G) G + j)
    std::vector<char> structured((size_t)std::max(NP, 1), 0);
    for (int p = 0; p < NP; ++p) {
        const int32_t *pi = d->phase_i + (size_t)p * PYXRD_PHASE_I_STRIDE;
        if (!(pi[4] & PYXRD_FLAG_VALID_PROBS) || (pi[4] & PYXRD_FLAG_RAW_PATTERN)) continue;
        const int G = pi[0], r = pi[1];
        if (G < 2 || r % G != 0) continue;
        const int model = (d->prob_model && d->prob_params) ? d->prob_model[p] : PYXRD_PROB_FIXED;
        if (model != PYXRD_PROB_FIXED) {
            structured[p] = (model == PYXRD_PROB_R2G2 || model == PYXRD_PROB_R3G2 || model == PYXRD_PROB_R2G3) ? 1
                            : model == PYXRD_PROB_R0 ? 2 : 0;
            continue;
        }
        const double *P = d->matrices + pi[6] + (size_t)r * r;
        if (r == G) {                                  // Reichweite 0: every row equals the first
            bool same = true;
            for (int I = 1; I < r && same; ++I)
                for (int J = 0; J < r; ++J)
                    if (P[(size_t)I * r + J] != P[J]) { same = false; break; }
            structured[p] = same ? 2 : 0;
            continue;
        }
        const int tails = r / G;
        bool ok = true;
        for (int I = 0; I < r && ok; ++I)
            for (int J = 0; J < r; ++J)
                if (J / G != I % tails && P[(size_t)I * r + J] != 0.0) { ok = false; break; }
        structured[p] = ok;
    }
    std::map<std::pair<int, int>, int> class_of;
    ctx->classes.clear();
    long long j = 0;
    for (int k = 0; k < K; ++k)
        for (int s = 0; s < S; ++s)
            for (int z = 0; z < Z; ++z)
                for (int m = 0; m < M; ++m, ++j) {
                    int pb = d->job_phase[j];
                    if (pb >= NP) return fail(ctx, "set_batch: job %lld references phase %d of %d", j, pb, NP);
                    job_spec[j] = s;
                    job_out[j] = (long long)k * cand_stride + (long long)M * Z * ctx->h_off[s] +
                                 ((long long)m * Z + z) * ctx->h_npts[s];
                    const int32_t *pi = pb >= 0 ? d->phase_i + (size_t)pb * PYXRD_PHASE_I_STRIDE : nullptr;
                    if (pb >= 0 && !(pi[4] & PYXRD_FLAG_VALID_PROBS)) pb = -1;    // invalid probabilities -> zeros
                    job_pb[j] = pb;
                    if (pb < 0 || ctx->h_npts[s] == 0) continue;
                    auto key = std::make_pair(pi[1] * 4 + structured[pb], pi[0]);
                    auto it = class_of.find(key);
                    if (it == class_of.end()) {
                        it = class_of.emplace(key, (int)ctx->classes.size()).first;
                        ctx->classes.emplace_back();
                        ctx->classes.back().rank = pi[1];
                        ctx->classes.back().G = pi[0];
                        ctx->classes.back().structure = structured[pb];
                    }
                    RankClass &rc = ctx->classes[it->second];
                    rc.jobs.push_back((int)j);
                    rc.max_x = std::max(rc.max_x, ctx->h_npts[s]);
                    rc.max_ntop_cap = std::max(rc.max_ntop_cap, pi[3] + 1);
                }
    // heavy classes first so the tail of the grid is filled by the cheap ones
    std::sort(ctx->classes.begin(), ctx->classes.end(), [](const RankClass &a, const RankClass &b) { return a.rank > b.rank; });
    jobids.clear();
    for (RankClass &rc : ctx->classes) jobids.insert(jobids.end(), rc.jobs.begin(), rc.jobs.end());
    return 0;
}

// derived buffers, the BatchDev view of the blob, and the per-candidate prologue kernels
int bind_batch(pyxrd_ctx *ctx, const BatchSizes &z, const BlobLayout &L, int ncap, int acap, bool keep_layer_tables = false) {
    const StaticDev &st = ctx->st;
    const int K = z.K, M = z.M, NP = z.NP, NC = z.NC, NA = z.NA, S = st.S, Z = st.Z;
    const long long cand_stride = (long long)M * Z * st.sumX;
    auto dev = [&](int idx) { return static_cast<char *>(ctx->d_blob.ptr) + L.off[idx]; };
    CK(ctx->d_atom_z.ensure(sizeof(double) * std::max(NA, 1)));
    CK(ctx->d_atom_plane.ensure(sizeof(int) * std::max(NA, 1)));
    CK(ctx->d_comp_np.ensure(sizeof(int) * std::max(NC, 1)));
    CK(ctx->d_comp_ilp.ensure(sizeof(int) * std::max(NC, 1)));
    {   // layer tables [n_classes][sumX] + the slack the multi-point threads of k_phase_small read past the end
        const size_t bytes = sizeof(double2) * ((size_t)std::max(z.n_classes, 1) * (size_t)st.sumX + PX_TABLE_PAD);
        if (bytes > ctx->d_layer_tab.cap) {
            CK(ctx->d_layer_tab.ensure(bytes));
            CK(cudaMemsetAsync(ctx->d_layer_tab.ptr, 0, ctx->d_layer_tab.cap, ctx->stream));
        }
    }
    CK(ctx->d_phase_x.ensure(sizeof(double) * 4 * std::max(NP, 1)));
    CK(ctx->d_phase_valid.ensure(sizeof(int) * std::max(NP, 1)));
    CK(ctx->d_coef.ensure(sizeof(double) * (size_t)ncap * std::max(NP, 1)));
    CK(ctx->d_ph_nu.ensure(sizeof(int) * std::max(NP, 1)));
    CK(ctx->d_ph_uz.ensure(sizeof(double) * (size_t)acap * std::max(NP, 1)));
    CK(ctx->d_ph_uend.ensure(sizeof(int) * (size_t)acap * std::max(NP, 1)));
    CK(ctx->d_ph_useg.ensure(sizeof(int) * (size_t)acap * std::max(NP, 1)));
    ctx->int_count = (int64_t)K * cand_stride;
    ctx->tot_count = (int64_t)K * Z * st.sumX;
    CK(ctx->d_int.ensure(sizeof(double) * std::max<int64_t>(ctx->int_count, 1)));
    ctx->wl_passes = 0;
    for (int s = 0; s < S; ++s)
        if (ctx->h_npts[s] > 0) ctx->wl_passes = std::max(ctx->wl_passes, ctx->h_nwl[s]);
    if (ctx->wl_passes > 0) CK(ctx->d_int2.ensure(sizeof(double) * std::max<int64_t>(ctx->int_count, 1)));
    ctx->jobs_per_cand = S * Z * M;
    CK(ctx->d_tot.ensure(sizeof(double) * std::max<int64_t>(ctx->tot_count, 1)));
    CK(ctx->d_res.ensure(sizeof(double) * (size_t)K * (S + 1)));
    CK(ctx->h_res.ensure(sizeof(double) * (size_t)K * (S + 1)));

    BatchDev &bt = ctx->bt;
    bt.K = K; bt.M = M; bt.NP = NP; bt.NC = NC; bt.NA = NA; bt.ncap = ncap;
    bt.cand_stride = cand_stride;
    bt.phase_i = reinterpret_cast<const int *>(dev(P_PHASE_I));
    bt.phase_d = reinterpret_cast<const double *>(dev(P_PHASE_D));
    bt.phase_comp = reinterpret_cast<const int *>(dev(P_PHASE_COMP));
    bt.matrices = reinterpret_cast<const double *>(dev(P_MATRICES));
    bt.comp_d = reinterpret_cast<const double *>(dev(P_COMP_D));
    bt.comp_i = reinterpret_cast<const int *>(dev(P_COMP_I));
    bt.atom_d = reinterpret_cast<const double *>(dev(P_ATOM_D));
    bt.atom_type = reinterpret_cast<const int *>(dev(P_ATOM_T));
    bt.atom_z = ctx->d_atom_z.as<double>();
    bt.atom_plane = ctx->d_atom_plane.as<int>();
    bt.comp_nplanes = ctx->d_comp_np.as<int>();
    bt.comp_ilplane = ctx->d_comp_ilp.as<int>();
    bt.comp_class = reinterpret_cast<const int *>(dev(P_COMP_CLASS));
    bt.class_rep = reinterpret_cast<const int *>(dev(P_CLASS_REP));
    bt.n_classes = z.n_classes;
    bt.layer_tab = ctx->d_layer_tab.as<double2>();
    bt.phase_x = ctx->d_phase_x.as<double>();
    bt.coef = ctx->d_coef.as<double>();
    bt.prob_model = z.model_driven ? reinterpret_cast<const int *>(dev(P_PMODEL)) : nullptr;
    bt.prob_params = z.model_driven ? reinterpret_cast<const double *>(dev(P_PPARAM)) : nullptr;
    bt.phase_valid = ctx->d_phase_valid.as<int>();
    bt.acap = acap;
    bt.ph_nu = ctx->d_ph_nu.as<int>();
    bt.ph_uz = ctx->d_ph_uz.as<double>();
    bt.ph_uend = ctx->d_ph_uend.as<int>();
    bt.ph_useg = ctx->d_ph_useg.as<int>();
    {
        // phase images (k_phase_image): capacities of this batch
        int mcap = 2;
        for (const RankClass &rc : ctx->classes) mcap = std::max(mcap, rc.rank * rc.rank + rc.G * rc.rank);
        bt.img_ae = (acap + 1) & ~1;
        bt.img_mcap = (mcap + 1) & ~1;
        bt.img_stride = PX_IMG_HDR + 4 * bt.img_ae + bt.img_mcap + ((bt.ncap + 1) & ~1);
        bt.img_ncopy = 0;
        CK(ctx->d_phase_img.ensure(sizeof(double) * (size_t)bt.img_stride * std::max(NP, 1)));
        bt.phase_img = ctx->d_phase_img.as<double>();
    }
    bt.job_pb = reinterpret_cast<const int *>(dev(P_JOB_PB));
    bt.job_spec = reinterpret_cast<const int *>(dev(P_JOB_SPEC));
    bt.job_out = reinterpret_cast<const long long *>(dev(P_JOB_OUT));
    bt.fractions = reinterpret_cast<const double *>(dev(P_FRAC));
    bt.scales = reinterpret_cast<const double *>(dev(P_SCALES));
    bt.bgshifts = reinterpret_cast<const double *>(dev(P_BG));
    bt.intensities = ctx->d_int.as<double>();
    bt.scaled = nullptr;
    bt.totals = ctx->d_tot.as<double>();
    bt.residuals = ctx->d_res.as<double>();
    ctx->NJ = (int)z.NJ;
    ctx->jobids_dev = reinterpret_cast<const int *>(dev(P_JOBIDS));
    {
        size_t pos = 0;
        for (RankClass &rc : ctx->classes) {
            rc.first = pos;
            pos += rc.count(K);
        }
    }

    CK(cudaEventRecord(ctx->ev[6], ctx->stream));
    // Two independent chains and one independent kernel, side by side (the kernels are small and latency-bound):
    //   probabilities -> phase prologue (CSDS, absolute scale, coefficient half of the phase images)   on the context's stream
    //   interlayer z-stretch -> distinct planes -> list half of the phase images (one kernel)          on an auxiliary stream
    //   layer tables (only when the classes or the static tables changed)  on another one
    cudaStream_t main_stream = ctx->stream;
    const bool side = ctx->overlap_classes && NP > 0;
    const bool need_tables = z.n_classes > 0 && st.sumX > 0 && !keep_layer_tables;
    if (side) {
        CK(cudaEventRecord(ctx->fork_ev, main_stream));
        CK(cudaStreamWaitEvent(ctx->class_stream[0], ctx->fork_ev, 0));
        if (need_tables) CK(cudaStreamWaitEvent(ctx->class_stream[1], ctx->fork_ev, 0));
    }
    if (NP > 0) {
        k_probabilities<<<(NP + 31) / 32, 32, 0, main_stream>>>(bt);        // serial per phase: spread over the SMs
        LAUNCH_CHECK("k_probabilities");
        k_phase_prologue<<<(NP + PX_PROLOGUE_WARPS - 1) / PX_PROLOGUE_WARPS, 32 * PX_PROLOGUE_WARPS, 0, main_stream>>>(bt);
        LAUNCH_CHECK("k_phase_prologue");
    }
    cudaStream_t z_stream = side ? ctx->class_stream[0] : main_stream;
    if (NP > 0) {
        // components referenced by several phases are stretched once, by a launch of their own
        const bool shared_components = z.NPC > z.NC;
        if (shared_components && NC > 0) {
            k_component_z<<<(NC + 63) / 64, 64, 0, z_stream>>>(bt);
            LAUNCH_CHECK("k_component_z");
        }
        k_phase_planes_image<<<NP, 32, 0, z_stream>>>(st, bt, shared_components ? 0 : 1);   // (z-stretch ->) distinct planes -> list half of the images
        LAUNCH_CHECK("k_phase_planes_image");
    }
    if (need_tables) {
        dim3 grid((unsigned)((st.sumX + 127) / 128), (unsigned)z.n_classes);
        k_layer_table<<<grid, 128, 0, side ? ctx->class_stream[1] : main_stream>>>(st, bt);
        LAUNCH_CHECK("k_layer_table");
    }
    if (side) {
        CK(cudaEventRecord(ctx->join_ev[0], ctx->class_stream[0]));
        CK(cudaStreamWaitEvent(main_stream, ctx->join_ev[0], 0));
        if (need_tables) {
            CK(cudaEventRecord(ctx->join_ev[1], ctx->class_stream[1]));
            CK(cudaStreamWaitEvent(main_stream, ctx->join_ev[1], 0));
        }
    }
    CK(cudaEventRecord(ctx->ev[7], ctx->stream));
    ctx->ev_set[3] = true;
    ctx->have_batch = true;
    return 0;
}

}  // namespace

extern "C" {

int pyxrd_set_batch(pyxrd_ctx *ctx, const pyxrd_batch_desc *d) {
    if (!ctx || !d) return 1;
    if (!ctx->have_static) return fail(ctx, "set_batch: call pyxrd_set_static first");
    CK(cudaSetDevice(ctx->device));
    ctx->have_batch = false;
    ctx->have_intensities = false;
    ctx->have_opt = false;
    ctx->have_totals = false;
    ctx->pop.valid = false;                  // the blob of a device-expanded population is overwritten
    const StaticDev &st = ctx->st;
    if (d->n_candidates <= 0 || d->n_phase_slots <= 0)
        return fail(ctx, "set_batch: need K > 0 and M > 0 (K=%d M=%d)", d->n_candidates, d->n_phase_slots);
    int ncap = 2, acap = 1;
    if (validate_batch(ctx, d, &ncap, &acap)) return 1;
    std::vector<int> job_pb, job_spec, jobids;
    std::vector<long long> job_out;
    if (build_jobs(ctx, d, job_pb, job_spec, job_out, jobids)) return 1;

    BatchSizes z;
    z.K = d->n_candidates; z.M = d->n_phase_slots; z.NP = d->n_phases; z.NC = d->n_components; z.NA = d->n_atoms;
    z.NPC = d->n_phase_comp; z.NM = d->n_matrix; z.NJ = (long long)job_pb.size(); z.njobids = jobids.size();
    z.model_driven = d->prob_model && d->prob_params;
    std::vector<int32_t> comp_class, class_rep;
    z.n_classes = layer_classes(d, nullptr, ctx->use_layer_tables ? 1 : 0, comp_class, class_rep);
    class_rep.resize(PX_MAX_LAYER_CLASSES, 0);
    const BlobLayout L = plan_blob(z, st.S);
    // ---- one pinned blob, one H2D copy ---------------------------------------------------
    const void *src[P_COUNT] = {d->phase_i, d->phase_d, d->phase_comp, d->matrices, d->comp_d, d->comp_i, d->atom_d, d->atom_type,
                                job_pb.data(), job_spec.data(), job_out.data(), jobids.data(), d->fractions, d->scales, d->bgshifts,
                                z.model_driven ? d->prob_model : nullptr, z.model_driven ? d->prob_params : nullptr,
                                comp_class.data(), class_rep.data()};
    CK(cudaStreamSynchronize(ctx->stream));        // the pinned blob may still feed a previous copy
    CK(ctx->h_blob.ensure(L.total));
    CK(ctx->d_blob.ensure(L.total));
    for (int i = 0; i < P_COUNT; ++i)
        if (L.bytes[i] && src[i]) memcpy(static_cast<char *>(ctx->h_blob.ptr) + L.off[i], src[i], L.bytes[i]);
    CK(cudaMemcpyAsync(ctx->d_blob.ptr, ctx->h_blob.ptr, L.total, cudaMemcpyHostToDevice, ctx->stream));
    return bind_batch(ctx, z, L, ncap, acap);
}


int pyxrd_set_solution(pyxrd_ctx *ctx, const double *fractions, const double *scales, const double *bgshifts) {
    if (!ctx || !ctx->have_batch) return fail(ctx, "set_solution: no resident batch");
    CK(cudaSetDevice(ctx->device));
    const BatchDev &bt = ctx->bt;
    const int S = ctx->st.S;
    CK(cudaMemcpyAsync(const_cast<double *>(bt.fractions), fractions, sizeof(double) * (size_t)bt.K * bt.M, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(const_cast<double *>(bt.scales), scales, sizeof(double) * (size_t)bt.K * S, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(const_cast<double *>(bt.bgshifts), bgshifts, sizeof(double) * (size_t)bt.K * S, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));     // sources are pageable caller memory
    return 0;
}

int pyxrd_set_intensities(pyxrd_ctx *ctx, const double *intensities, int64_t count) {
    if (!ctx || !ctx->have_batch) return fail(ctx, "set_intensities: no resident batch");
    if (count != ctx->int_count) return fail(ctx, "set_intensities: got %lld doubles, the batch holds %lld", (long long)count, (long long)ctx->int_count);
    CK(cudaSetDevice(ctx->device));
    if (count) CK(cudaMemcpyAsync(ctx->d_int.ptr, intensities, sizeof(double) * (size_t)count, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_intensities = true;
    return 0;
}

}  // extern "C"

namespace {

// phase patterns + wavelength passes of the candidates [k0, k1) on ctx->stream.  The phase kernels
// write the buffer from which ctx->wl_passes out-of-place passes end in d_int.
// phase kernels of the candidates [k0, k1) into `target` (one of the two pattern buffers)
int phase_range(pyxrd_ctx *ctx, int k0, int k1, double *target, bool timed) {
    BatchDev bt = ctx->bt;
    bt.intensities = target;
    const long long j0 = (long long)k0 * ctx->jobs_per_cand, j1 = (long long)k1 * ctx->jobs_per_cand;
    // None phases / invalid probabilities / empty specimens stay zero (specimen.py:63, phases.py:109-111)
    size_t listed = 0;
    for (const RankClass &rc : ctx->classes) listed += rc.count(ctx->bt.K);
    if (listed != (size_t)ctx->NJ && bt.cand_stride > 0)
        CK(cudaMemsetAsync(bt.intensities + (long long)k0 * bt.cand_stride, 0,
                           sizeof(double) * (size_t)(k1 - k0) * bt.cand_stride, ctx->stream));
    if (timed) CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    // the heaviest class on the context's stream, the others beside it on auxiliary streams: the classes write
    // disjoint patterns, and the tail of one launch is filled by the blocks of the next
    cudaStream_t main_stream = ctx->stream;
    const bool overlap = ctx->overlap_classes && ctx->classes.size() > 1;
    if (overlap) CK(cudaEventRecord(ctx->fork_ev, main_stream));
    bool used[3] = {false, false, false};
    int index = 0, rc_status = 0;
    for (const RankClass &rc : ctx->classes) {
        size_t a, b;
        rc.range(k0, k1, j0, j1, a, b);
        if (b <= a) continue;
        const int aux = (index - 1) % 3;
        if (overlap && index > 0) {
            if (!used[aux]) CK(cudaStreamWaitEvent(ctx->class_stream[aux], ctx->fork_ev, 0));
            used[aux] = true;
            ctx->stream = ctx->class_stream[aux];
        }
        rc_status = launch_class(ctx, rc, ctx->jobids_dev + rc.first + a, b - a, bt);
        ctx->stream = main_stream;
        if (rc_status) return 1;
        ++index;
    }
    for (int i = 0; i < 3; ++i)
        if (used[i]) {
            CK(cudaEventRecord(ctx->join_ev[i], ctx->class_stream[i]));
            CK(cudaStreamWaitEvent(main_stream, ctx->join_ev[i], 0));
        }
    if (timed) {
        CK(cudaEventRecord(ctx->ev[1], ctx->stream));
        CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    }
    return 0;
}

// phase patterns + wavelength passes of the candidates [k0, k1).  fuse_last: the last pass of the wavelength
// distribution is left to the residual kernel (k_mixture_residual<true>); *last_src then names the buffer it reads.
int residuals_range(pyxrd_ctx *ctx, int k0, int k1, int flags, bool timed, const double *gather_src, bool store_patterns, bool all_passes,
                    bool patterns_only);
bool patterns_fusable(const pyxrd_ctx *ctx) {
    return ctx->fuse_patterns && ctx->wl_passes >= 1 && ctx->wl_passes <= PX_FUSE_MAXPASS && ctx->wl_halo_l + ctx->wl_halo_r <= PX_FUSE_HMAX;
}

int intensities_range(pyxrd_ctx *ctx, int k0, int k1, bool timed, bool fuse_last = false, const double **last_src = nullptr) {
    const StaticDev &st = ctx->st;
    double *bufs[2] = {ctx->d_int.as<double>(), ctx->d_int2.as<double>()};
    if (!fuse_last && ctx->wl_passes >= 2 && patterns_fusable(ctx)) {
        // all passes in ONE kernel (k_patterns_residual storing the patterns; its tile sums are scratch): one read and
        // one write of the patterns instead of one per pass
        if (phase_range(ctx, k0, k1, bufs[1], timed)) return 1;
        if (residuals_range(ctx, k0, k1, PYXRD_RESIDUALS_ONLY, false, bufs[1], true, true, true)) return 1;
        if (last_src) *last_src = bufs[0];
        if (timed) {
            CK(cudaEventRecord(ctx->ev[3], ctx->stream));
            ctx->ev_set[0] = ctx->ev_set[1] = true;
        }
        return 0;
    }
    int cur = ctx->wl_passes & 1;                      // passes alternate cur -> 1 - cur and must end in 0
    const long long j0 = (long long)k0 * ctx->jobs_per_cand, j1 = (long long)k1 * ctx->jobs_per_cand;
    if (phase_range(ctx, k0, k1, bufs[cur], timed)) return 1;
    const int separate = ctx->wl_passes - (fuse_last && ctx->wl_passes > 0 ? 1 : 0);
    if (j1 > j0 && st.maxX > 0)
        for (int pass = 0; pass < separate; ++pass) {
            dim3 grid((unsigned)(j1 - j0), (unsigned)((st.maxX + PX_WL_TPB * PX_WL_PPT - 1) / (PX_WL_TPB * PX_WL_PPT)));
            k_wavelength_pass<<<grid, PX_WL_TPB, 0, ctx->stream>>>(st, ctx->bt, (int)j0, pass, bufs[cur], bufs[1 - cur]);
            LAUNCH_CHECK("k_wavelength_pass");
            cur = 1 - cur;
        }
    if (last_src) *last_src = bufs[cur];
    if (timed) {
        CK(cudaEventRecord(ctx->ev[3], ctx->stream));
        ctx->ev_set[0] = ctx->ev_set[1] = true;
    }
    return 0;
}

// mixture totals + residuals of the candidates [k0, k1).  flags: bit 0 = also the scaled phase patterns, bit 1 =
// residuals only (the total patterns are summed but not stored).  gather_src != nullptr: the last wavelength pass is
// taken on the fly from that buffer and the final patterns are written to d_int by the same kernel.
int residuals_range(pyxrd_ctx *ctx, int k0, int k1, int flags, bool timed, const double *gather_src = nullptr,
                    bool store_patterns = true, bool all_passes = false, bool patterns_only = false) {
    BatchDev bt = ctx->bt;
    const int want_scaled = flags & 1;
    int want_totals = (flags & 2) ? 0 : 1;
    if (ctx->residual_method == PYXRD_RESIDUAL_RPDER) want_totals = 1;      // k_residual_rpder reads the totals
    if (patterns_only) want_totals = 0;
    if (want_scaled) {
        CK(ctx->d_scaled.ensure(sizeof(double) * std::max<int64_t>(ctx->int_count, 1)));
        bt.scaled = ctx->d_scaled.as<double>();
    }
    if (timed) CK(cudaEventRecord(ctx->ev[4], ctx->stream));
    if (k1 > k0) {
        dim3 grid(k1 - k0, ctx->st.S);
        const int tile = PX_RES_TPB * PX_RES_PPT, ntiles = std::max(1, (ctx->st.maxX + tile - 1) / tile);
        CK(ctx->d_partial.ensure(sizeof(double) * (size_t)(k1 - k0) * ctx->st.S * ntiles));
        dim3 tgrid(k1 - k0, ntiles, ctx->st.S);
        if (all_passes) {
            static_assert(PX_FUSE_T == PX_RES_TPB * PX_RES_PPT, "the two residual kernels share the tile sums");
            double *dst = store_patterns ? ctx->d_int.as<double>() : nullptr;
            const int HL = ctx->wl_halo_l, HR = ctx->wl_halo_r;
            // candidates per block: up to PX_FUSE_CPB, fewer where that would leave SMs without blocks (two waves of 4)
            const long long tiles_total = (long long)(k1 - k0) * ntiles * ctx->st.S;
            const int cpb = (int)std::max<long long>(1, std::min<long long>(PX_FUSE_CPB, tiles_total / (8LL * ctx->sm_count)));
            dim3 fgrid((k1 - k0 + cpb - 1) / cpb, ntiles, ctx->st.S);
            const bool stores = dst || want_scaled || want_totals;
#define PX_FUSED(N_) do { if (stores) k_patterns_residual<N_, true><<<fgrid, PX_FUSE_TPB, 0, ctx->stream>>>(ctx->st, bt, k0, k1, cpb, want_scaled, want_totals, HL, HR, gather_src, dst, ctx->d_partial.as<double>()); \
                          else k_patterns_residual<N_, false><<<fgrid, PX_FUSE_TPB, 0, ctx->stream>>>(ctx->st, bt, k0, k1, cpb, 0, 0, HL, HR, gather_src, nullptr, ctx->d_partial.as<double>()); } while (0)
            switch (ctx->wl_passes) {
                case 1: PX_FUSED(1); break;
                case 2: PX_FUSED(2); break;
                case 3: PX_FUSED(3); break;
                default: PX_FUSED(4); break;
            }
#undef PX_FUSED
        } else if (gather_src)
            k_mixture_residual<true><<<tgrid, PX_RES_TPB, 0, ctx->stream>>>(ctx->st, bt, k0, want_scaled, want_totals, ctx->wl_passes - 1,
                                                                            gather_src, ctx->d_int.as<double>(), ctx->d_partial.as<double>());
        else
            k_mixture_residual<false><<<tgrid, PX_RES_TPB, 0, ctx->stream>>>(ctx->st, bt, k0, want_scaled, want_totals, 0,
                                                                             ctx->d_int.as<double>(), nullptr, ctx->d_partial.as<double>());
        LAUNCH_CHECK("k_mixture_residual");
        if (patterns_only) return 0;            // (intensities_range: the kernel was run for the patterns it stores)
        if (ctx->residual_method == PYXRD_RESIDUAL_RP) {
            k_residual_finish_mean<<<(k1 - k0 + 127) / 128, 128, 0, ctx->stream>>>(ctx->st, bt, k0, k1, ntiles, ctx->d_partial.as<double>());
            LAUNCH_CHECK("k_residual_finish_mean");
            ctx->have_totals = want_totals != 0;
            if (timed) {
                CK(cudaEventRecord(ctx->ev[5], ctx->stream));
                ctx->ev_set[2] = true;
            }
            return 0;
        }
        k_residual_finish<<<((k1 - k0) * ctx->st.S + 127) / 128, 128, 0, ctx->stream>>>(ctx->st, bt, k0, k1, ntiles,
                                                                                      ctx->d_partial.as<double>());
        LAUNCH_CHECK("k_residual_finish");
        if (ctx->residual_method == PYXRD_RESIDUAL_RPDER) {
            // the Rp kernel above has produced the totals; its residuals are replaced
            int maxn = 0;
            for (int s = 0; s < ctx->st.S; ++s) {
                const int n = ctx->st.Z * ctx->h_selcount[s];
                if (ctx->h_npts[s] > 0 && n < 31)
                    return fail(ctx, "Input vector needs to be bigger than window size.");   // math_tools.py:87-88
                maxn = std::max(maxn, n);
            }
            const size_t smem = sizeof(double) * 2 * (size_t)maxn;
            if (smem + 1024 <= (size_t)ctx->smem_optin) {
                if (set_smem(ctx, k_residual_rpder<true>, smem)) return 1;
                k_residual_rpder<true><<<grid, 256, smem, ctx->stream>>>(ctx->st, bt, k0, ctx->d_selpos.as<int>(), ctx->d_selcount.as<int>());
            } else {
                k_residual_rpder<false><<<grid, 256, 0, ctx->stream>>>(ctx->st, bt, k0, ctx->d_selpos.as<int>(), ctx->d_selcount.as<int>());
            }
            LAUNCH_CHECK("k_residual_rpder");
        }
        k_mixture_mean<<<(k1 - k0 + 127) / 128, 128, 0, ctx->stream>>>(k0, k1, ctx->st.S, bt.residuals);
        LAUNCH_CHECK("k_mixture_mean");
    }
    ctx->have_totals = want_totals != 0;
    if (timed) {
        CK(cudaEventRecord(ctx->ev[5], ctx->stream));
        ctx->ev_set[2] = true;
    }
    return 0;
}

// patterns + residuals of the candidates [k0, k1) at the batch's solution: phase kernels, all wavelength passes but the
// last as streaming kernels, the last one inside the residual kernel.  (A single pass that composes ALL the wavelength
// entries per point was measured in round 1: three levels of dependent gathers per point, slower than the passes.)
int evaluate_range(pyxrd_ctx *ctx, int k0, int k1, int flags, bool timed, bool store_patterns = true) {
    const double *src = nullptr;
    const bool fused = patterns_fusable(ctx);
    if (fused) {
        // raw phase patterns into the second buffer; k_patterns_residual takes every pass of the distribution from there
        if (phase_range(ctx, k0, k1, ctx->d_int2.as<double>(), timed)) return 1;
        if (timed) {
            CK(cudaEventRecord(ctx->ev[3], ctx->stream));
            ctx->ev_set[0] = ctx->ev_set[1] = true;
        }
        return residuals_range(ctx, k0, k1, flags, timed, ctx->d_int2.as<double>(), store_patterns, true);
    }
    if (intensities_range(ctx, k0, k1, timed, true, &src)) return 1;
    return residuals_range(ctx, k0, k1, flags, timed, ctx->wl_passes > 0 ? src : nullptr);
}

}  // namespace

extern "C" {

int pyxrd_compute_all(pyxrd_ctx *ctx, int flags) {
    if (!ctx || !ctx->have_batch) return fail(ctx, "compute_all: no resident batch");
    CK(cudaSetDevice(ctx->device));
    // residuals only (no totals, no scaled patterns, Rp): nobody reads the final phase patterns; the fused kernel does not
    // store them and the pattern buffer is marked as not computed
    const bool store = !((flags & PYXRD_RESIDUALS_ONLY) && !(flags & PYXRD_WANT_SCALED) && ctx->residual_method == PYXRD_RESIDUAL_RP);
    if (evaluate_range(ctx, 0, ctx->bt.K, flags, true, store)) return 1;
    ctx->have_intensities = store || !patterns_fusable(ctx);
    return 0;
}

int pyxrd_compute_intensities(pyxrd_ctx *ctx) {
    if (!ctx || !ctx->have_batch) return fail(ctx, "compute_intensities: no resident batch");
    CK(cudaSetDevice(ctx->device));
    if (intensities_range(ctx, 0, ctx->bt.K, true)) return 1;
    ctx->have_intensities = true;
    return 0;
}

int pyxrd_set_residual_method(pyxrd_ctx *ctx, int method) {
    if (!ctx) return 1;
    if (method != PYXRD_RESIDUAL_RP && method != PYXRD_RESIDUAL_RPDER)
        return fail(ctx, "set_residual_method: unknown method %d (0 = Rp, 1 = Rpder)", method);
    ctx->residual_method = method;
    return 0;
}

int pyxrd_compute_residuals(pyxrd_ctx *ctx, int flags) {
    if (!ctx || !ctx->have_batch || !ctx->have_intensities) return fail(ctx, "compute_residuals: intensities not computed");
    CK(cudaSetDevice(ctx->device));
    return residuals_range(ctx, 0, ctx->bt.K, flags, true);
}

int pyxrd_optimize(pyxrd_ctx *ctx, int refine_bg, int outer_iterations, int newton_iterations) {
    if (!ctx || !ctx->have_batch || !ctx->have_intensities) return fail(ctx, "optimize: intensities not computed");
    CK(cudaSetDevice(ctx->device));
    const int M = ctx->bt.M, S = ctx->st.S, K = ctx->bt.K;
    if (M + 1 > PX_OPT_MAXMB || S > PX_OPT_MAXS)
        return fail(ctx, "optimize: the device optimiser handles up to %d phases and %d specimens (got %d, %d)",
                    PX_OPT_MAXMB - 1, PX_OPT_MAXS, M, S);
    CK(ctx->d_sol.ensure(sizeof(double) * (size_t)K * (M + 2 * S)));
    CK(ctx->d_optres.ensure(sizeof(double) * (size_t)K * (S + 1)));
    OptParams op;
    op.outer = outer_iterations > 0 ? outer_iterations : 32;
    op.lm_iters = newton_iterations > 0 ? newton_iterations : 2;
    op.refine_bg = refine_bg ? 1 : 0;
    op.single_from = ctx->opt_single_from;
    op.newton_tol = ctx->opt_newton_tol;
    op.lane_solver = ctx->opt_lane_solver;
    op.coarse_until = std::min(ctx->opt_coarse_until, op.outer);
    op.coarse_stride = ctx->opt_coarse_stride;
    op.delta0 = ctx->opt_delta0;
    op.shrink = ctx->opt_shrink;
    op.delta_min = ctx->opt_delta_min;
    CK(cudaEventRecord(ctx->ev[4], ctx->stream));
    int rc = 1;
    switch (M + 1) {
        case 2: rc = launch_optimize<2>(ctx, op); break;
        case 3: rc = launch_optimize<3>(ctx, op); break;
        case 4: rc = launch_optimize<4>(ctx, op); break;
        case 5: rc = launch_optimize<5>(ctx, op); break;
        case 6: rc = launch_optimize<6>(ctx, op); break;
        case 7: rc = launch_optimize<7>(ctx, op); break;
        case 8: rc = launch_optimize<8>(ctx, op); break;
        case 9: rc = launch_optimize<9>(ctx, op); break;
        case 10: rc = launch_optimize<10>(ctx, op); break;
        case 11: rc = launch_optimize<11>(ctx, op); break;
        case 12: rc = launch_optimize<12>(ctx, op); break;
        default: return fail(ctx, "optimize: unsupported phase count %d", M);
    }
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev[5], ctx->stream));
    ctx->ev_set[2] = true;
    ctx->have_opt = true;
    return 0;
}

int pyxrd_read_solutions(pyxrd_ctx *ctx, double *out, int64_t count) {
    if (!ctx || !ctx->have_opt) return fail(ctx, "read solutions: optimiser has not run");
    return read_back(ctx, ctx->d_sol.ptr, out, count, (int64_t)ctx->bt.K * (ctx->bt.M + 2 * ctx->st.S), "solutions");
}
int pyxrd_read_opt_residuals(pyxrd_ctx *ctx, double *out, int64_t count) {
    if (!ctx || !ctx->have_opt) return fail(ctx, "read opt residuals: optimiser has not run");
    return read_back(ctx, ctx->d_optres.ptr, out, count, (int64_t)ctx->bt.K * (ctx->st.S + 1), "optimised residuals");
}

int64_t pyxrd_intensity_count(const pyxrd_ctx *ctx) { return ctx ? ctx->int_count : 0; }
int64_t pyxrd_total_count(const pyxrd_ctx *ctx) { return ctx ? ctx->tot_count : 0; }

int pyxrd_read_intensities(pyxrd_ctx *ctx, double *out, int64_t count) {
    if (!ctx || !ctx->have_intensities) return fail(ctx, "read intensities: not computed");
    return read_back(ctx, ctx->d_int.ptr, out, count, ctx->int_count, "intensities");
}
int pyxrd_read_scaled(pyxrd_ctx *ctx, double *out, int64_t count) {
    if (!ctx || !ctx->d_scaled.ptr) return fail(ctx, "read scaled: not computed");
    return read_back(ctx, ctx->d_scaled.ptr, out, count, ctx->int_count, "scaled intensities");
}
int pyxrd_read_totals(pyxrd_ctx *ctx, double *out, int64_t count) {
    if (!ctx || !ctx->have_batch) return fail(ctx, "read totals: no batch");
    if (!ctx->have_totals) return fail(ctx, "read totals: the last residual call did not store them (PYXRD_RESIDUALS_ONLY)");
    return read_back(ctx, ctx->d_tot.ptr, out, count, ctx->tot_count, "totals");
}
int pyxrd_read_residuals(pyxrd_ctx *ctx, double *out, int64_t count) {
    if (!ctx || !ctx->have_batch) return fail(ctx, "read residuals: no batch");
    return read_back(ctx, ctx->d_res.ptr, out, count, (int64_t)ctx->bt.K * (ctx->st.S + 1), "residuals");
}
int pyxrd_read_correction(pyxrd_ctx *ctx, double *out, int64_t count) {
    if (!ctx || !ctx->have_static) return fail(ctx, "read correction: no static tables");
    return read_back(ctx, ctx->d_corr.ptr, out, count, ctx->st.sumX, "correction");
}

int pyxrd_device_buffers(pyxrd_ctx *ctx, double **intensities_dev, double **totals_dev, double **residuals_dev) {
    if (!ctx || !ctx->have_batch) return fail(ctx, "device_buffers: no batch");
    if (intensities_dev) *intensities_dev = ctx->d_int.as<double>();
    if (totals_dev) *totals_dev = ctx->d_tot.as<double>();
    if (residuals_dev) *residuals_dev = ctx->d_res.as<double>();
    return 0;
}

int pyxrd_device_results(pyxrd_ctx *ctx, double **opt_residuals_dev, double **solutions_dev) {
    if (!ctx || !ctx->have_opt) return fail(ctx, "device_results: optimiser has not run");
    if (opt_residuals_dev) *opt_residuals_dev = ctx->d_optres.as<double>();
    if (solutions_dev) *solutions_dev = ctx->d_sol.as<double>();
    return 0;
}

int pyxrd_put_rows_to_peers(pyxrd_ctx *ctx, const double *src_dev, int64_t n_rows, int cols, void *const *peer_tables,
                            int n_peers, int64_t dst_row0) {
    if (!ctx) return 1;
    if (!src_dev || !peer_tables || n_rows < 0 || cols <= 0 || n_peers <= 0 || n_peers > PX_MAX_PEERS || dst_row0 < 0)
        return fail(ctx, "put_rows_to_peers: bad arguments (rows %lld, cols %d, peers %d of at most %d)", (long long)n_rows, cols, n_peers, PX_MAX_PEERS);
    CK(cudaSetDevice(ctx->device));
    if (n_rows == 0) return 0;
    PeerTables peers{};
    peers.n = n_peers;
    for (int p = 0; p < n_peers; ++p) {
        if (!peer_tables[p]) return fail(ctx, "put_rows_to_peers: peer %d has no table", p);
        peers.table[p] = static_cast<double *>(peer_tables[p]);
    }
    const long long count = (long long)n_rows * cols;
    k_put_rows_to_peers<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(src_dev, count, (long long)dst_row0 * cols, peers);
    LAUNCH_CHECK("k_put_rows_to_peers");
    return 0;
}

int pyxrd_evaluate_host(pyxrd_ctx *ctx, const pyxrd_batch_desc *desc, double *residuals_out, double *intensities_out) {
    if (pyxrd_set_batch(ctx, desc)) return 1;
    const int K = ctx->bt.K;
    const size_t nres = (size_t)K * (ctx->st.S + 1);
    // candidate chunks: while chunk c + 1 is computed, the patterns of chunk c travel to the host on
    // the copy stream (the patterns are the bulk of the result: 8 B per point against ~2 kflop)
    int nchunks = 1;
    if (intensities_out && K >= 16) nchunks = std::min(8, K / 8);
    if (ctx->host_chunks > 0) nchunks = std::max(1, std::min(16, std::min(K, ctx->host_chunks)));
    for (int c = 0; c < nchunks; ++c) {
        const int k0 = (int)((long long)K * c / nchunks), k1 = (int)((long long)K * (c + 1) / nchunks);
        if (evaluate_range(ctx, k0, k1, PYXRD_RESIDUALS_ONLY, nchunks == 1)) return 1;
        if (intensities_out && ctx->bt.cand_stride > 0) {
            CK(cudaEventRecord(ctx->chunk_ev[c], ctx->stream));
            CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->chunk_ev[c], 0));
            CK(cudaMemcpyAsync(intensities_out + (long long)k0 * ctx->bt.cand_stride,
                               ctx->d_int.as<double>() + (long long)k0 * ctx->bt.cand_stride,
                               sizeof(double) * (size_t)(k1 - k0) * ctx->bt.cand_stride, cudaMemcpyDeviceToHost, ctx->copy_stream));
        }
    }
    ctx->have_intensities = true;
    CK(cudaMemcpyAsync(ctx->h_res.ptr, ctx->d_res.ptr, sizeof(double) * nres, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->copy_stream));
    memcpy(residuals_out, ctx->h_res.ptr, sizeof(double) * nres);
    return 0;
}

int pyxrd_optimize_host(pyxrd_ctx *ctx, const pyxrd_batch_desc *desc, int refine_bg, int outer_iterations,
                        int newton_iterations, double *residuals_out, double *solutions_out) {
    if (pyxrd_set_batch(ctx, desc)) return 1;
    if (pyxrd_compute_intensities(ctx)) return 1;
    if (pyxrd_optimize(ctx, refine_bg, outer_iterations, newton_iterations)) return 1;
    const size_t nres = (size_t)ctx->bt.K * (ctx->st.S + 1);
    CK(cudaMemcpyAsync(ctx->h_res.ptr, ctx->d_optres.ptr, sizeof(double) * nres, cudaMemcpyDeviceToHost, ctx->stream));
    if (solutions_out)
        CK(cudaMemcpyAsync(solutions_out, ctx->d_sol.ptr, sizeof(double) * (size_t)ctx->bt.K * (ctx->bt.M + 2 * ctx->st.S),
                           cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    memcpy(residuals_out, ctx->h_res.ptr, sizeof(double) * nres);
    return 0;
}

// ---- leaf helpers -----------------------------------------------------------------------
int pyxrd_leaf_asf(pyxrd_ctx *ctx, const double *s, int64_t n, const double *atom_type, double *out) {
    if (!ctx || n < 0) return 1;
    CK(cudaSetDevice(ctx->device));
    if (n == 0) return 0;
    if (upload(ctx, ctx->d_leaf_a, s, (size_t)n)) return 1;
    if (upload(ctx, ctx->d_leaf_b, atom_type, PYXRD_ATYPE_STRIDE)) return 1;
    CK(ctx->d_leaf_c.ensure(sizeof(double) * (size_t)n));
    k_leaf_asf<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_leaf_a.as<double>(), n, ctx->d_leaf_b.as<double>(), ctx->d_leaf_c.as<double>());
    LAUNCH_CHECK("k_leaf_asf");
    return read_back(ctx, ctx->d_leaf_c.ptr, out, n, n, "asf");
}

int pyxrd_leaf_statistic(pyxrd_ctx *ctx, int kind, const double *exp, const double *calc, const double *phase, int64_t n,
                         double extra, double *out) {
    if (!ctx || !exp || !out || n < 0) return 1;
    if (kind < 0 || kind > 4) return fail(ctx, "leaf_statistic: unknown statistic %d", kind);
    if ((kind != 3 && !calc) || (kind == 4 && !phase)) return fail(ctx, "leaf_statistic: statistic %d needs more patterns", kind);
    CK(cudaSetDevice(ctx->device));
    if (upload(ctx, ctx->d_leaf_a, exp, (size_t)n)) return 1;
    if (calc && upload(ctx, ctx->d_leaf_b, calc, (size_t)n)) return 1;
    if (kind == 4 && upload(ctx, ctx->d_leaf_d, phase, (size_t)n)) return 1;
    CK(ctx->d_leaf_c.ensure(sizeof(double) * 2));
    k_leaf_statistic<<<1, 256, 0, ctx->stream>>>(kind, ctx->d_leaf_a.as<double>(), calc ? ctx->d_leaf_b.as<double>() : nullptr,
                                                 kind == 4 ? ctx->d_leaf_d.as<double>() : nullptr, (long long)n, extra,
                                                 ctx->d_leaf_c.as<double>());
    LAUNCH_CHECK("k_leaf_statistic");
    return read_back(ctx, ctx->d_leaf_c.ptr, out, 1, 1, "statistic");
}

int pyxrd_leaf_derive(pyxrd_ctx *ctx, const double *pattern, int64_t n, int smooth_only, double *out) {
    if (!ctx || !pattern || !out || n < 0) return 1;
    if (n < 31) return fail(ctx, "Input vector needs to be bigger than window size.");       // math_tools.py:87-88
    if (n > 0x7fffffffLL) return fail(ctx, "leaf_derive: pattern too long");
    CK(cudaSetDevice(ctx->device));
    if (upload(ctx, ctx->d_leaf_a, pattern, (size_t)n)) return 1;
    CK(ctx->d_leaf_c.ensure(sizeof(double) * (size_t)n));
    k_leaf_derive<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_leaf_a.as<double>(), (int)n, smooth_only,
                                                                      ctx->d_leaf_c.as<double>());
    LAUNCH_CHECK("k_leaf_derive");
    return read_back(ctx, ctx->d_leaf_c.ptr, out, n, n, "derive");
}

int pyxrd_leaf_lpf(pyxrd_ctx *ctx, const double *theta, int64_t n, double sigma_star, double soller1, double soller2,
                   double mcr_2theta, double *out) {
    if (!ctx || n < 0) return 1;
    CK(cudaSetDevice(ctx->device));
    if (n == 0) return 0;
    if (upload(ctx, ctx->d_leaf_a, theta, (size_t)n)) return 1;
    CK(ctx->d_leaf_c.ensure(sizeof(double) * (size_t)n));
    k_leaf_lpf<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_leaf_a.as<double>(), n, sigma_star, soller1, soller2, mcr_2theta, ctx->d_leaf_c.as<double>(), 0);
    LAUNCH_CHECK("k_leaf_lpf");
    return read_back(ctx, ctx->d_leaf_c.ptr, out, n, n, "lpf");
}

int pyxrd_leaf_lorentz(pyxrd_ctx *ctx, const double *theta, int64_t n, double sigma_star, double soller1, double soller2, double *out) {
    if (!ctx || n < 0) return 1;
    CK(cudaSetDevice(ctx->device));
    if (n == 0) return 0;
    if (upload(ctx, ctx->d_leaf_a, theta, (size_t)n)) return 1;
    CK(ctx->d_leaf_c.ensure(sizeof(double) * (size_t)n));
    k_leaf_lpf<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_leaf_a.as<double>(), n, sigma_star, soller1, soller2, 0.0, ctx->d_leaf_c.as<double>(), 1);
    LAUNCH_CHECK("k_leaf_lpf");
    return read_back(ctx, ctx->d_leaf_c.ptr, out, n, n, "T");
}

int pyxrd_leaf_csds(pyxrd_ctx *ctx, const double *csds6, int minimum, int maximum, double *out_arr, double *out_mean, double *out_coef) {
    if (!ctx) return 1;
    if (minimum < 0 || maximum < minimum) return fail(ctx, "leaf_csds: unsupported range [%d, %d]", minimum, maximum);
    CK(cudaSetDevice(ctx->device));
    double pd[8] = {0.0, csds6[0], csds6[1], csds6[2], csds6[3], csds6[4], 0.0, 0.0};
    if (upload(ctx, ctx->d_leaf_a, pd, 8)) return 1;
    const size_t n_arr = (size_t)maximum + 1, n_coef = (size_t)maximum + 2;
    CK(ctx->d_leaf_c.ensure(sizeof(double) * (n_arr + 1 + n_coef)));
    double *base = ctx->d_leaf_c.as<double>();
    k_leaf_csds<<<1, 32, 0, ctx->stream>>>(ctx->d_leaf_a.as<double>(), minimum, maximum, base, base + n_arr, base + n_arr + 1);
    LAUNCH_CHECK("k_leaf_csds");
    std::vector<double> host(n_arr + 1 + n_coef);
    if (read_back(ctx, base, host.data(), (int64_t)host.size(), (int64_t)host.size(), "csds")) return 1;
    memcpy(out_arr, host.data(), sizeof(double) * n_arr);
    *out_mean = host[n_arr];
    memcpy(out_coef, host.data() + n_arr + 1, sizeof(double) * n_coef);
    return 0;
}

int pyxrd_leaf_component(pyxrd_ctx *ctx, const double *stl, int64_t n, const double *comp_d, int n_layer, int n_interlayer,
                         const double *atom_d, const double *atom_type_rows, const uint8_t *atom_has_type,
                         double *sf_out, double *phi_out, double *z_out) {
    if (!ctx || n < 0 || n_layer < 0 || n_interlayer < 0) return 1;
    CK(cudaSetDevice(ctx->device));
    const int NA = n_layer + n_interlayer;
    const size_t npt = (size_t)std::max<int64_t>(n, 1);
    // one scratch allocation: stl | comp | atom_d | rows | has | sf | phi | z
    size_t off_stl = 0, off_comp = align_up(off_stl + 8 * npt, 256), off_atom = align_up(off_comp + 64, 256);
    size_t off_rows = align_up(off_atom + 16 * (size_t)std::max(NA, 1), 256);
    size_t off_has = align_up(off_rows + 96 * (size_t)std::max(NA, 1), 256);
    size_t off_sf = align_up(off_has + (size_t)std::max(NA, 1), 256), off_phi = align_up(off_sf + 16 * npt, 256);
    size_t off_z = align_up(off_phi + 16 * npt, 256), total = off_z + 8 * (size_t)std::max(NA, 1);
    CK(ctx->d_leaf_a.ensure(total));
    char *base = ctx->d_leaf_a.as<char>();
    if (n) CK(cudaMemcpyAsync(base + off_stl, stl, 8 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(base + off_comp, comp_d, 64, cudaMemcpyHostToDevice, ctx->stream));
    if (NA) {
        CK(cudaMemcpyAsync(base + off_atom, atom_d, 16 * (size_t)NA, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(base + off_rows, atom_type_rows, 96 * (size_t)NA, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(base + off_has, atom_has_type, (size_t)NA, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (n) {
        k_leaf_component<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(
            reinterpret_cast<double *>(base + off_stl), n, reinterpret_cast<double *>(base + off_comp), n_layer, n_interlayer,
            reinterpret_cast<double *>(base + off_atom), reinterpret_cast<double *>(base + off_rows),
            reinterpret_cast<unsigned char *>(base + off_has), reinterpret_cast<double2 *>(base + off_sf),
            reinterpret_cast<double2 *>(base + off_phi), reinterpret_cast<double *>(base + off_z));
        LAUNCH_CHECK("k_leaf_component");
        CK(cudaMemcpyAsync(sf_out, base + off_sf, 16 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(phi_out, base + off_phi, 16 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        if (NA) CK(cudaMemcpyAsync(z_out, base + off_z, 8 * (size_t)NA, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int pyxrd_leaf_mmult(pyxrd_ctx *ctx, const double *a, const double *b, int64_t n_points, int rank, double *out) {
    if (!ctx || !a || !b || !out || n_points < 0 || rank <= 0) return 1;
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)n_points * rank * rank * 2;
    if (n == 0) return 0;
    if (upload(ctx, ctx->d_leaf_a, a, n)) return 1;
    if (upload(ctx, ctx->d_leaf_b, b, n)) return 1;
    CK(ctx->d_leaf_c.ensure(sizeof(double) * n));
    const long long threads = (long long)n_points * rank * rank;
    k_leaf_mmult<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_leaf_a.as<double2>(), ctx->d_leaf_b.as<double2>(),
                                                                           (long long)n_points, rank, ctx->d_leaf_c.as<double2>());
    LAUNCH_CHECK("k_leaf_mmult");
    return read_back(ctx, ctx->d_leaf_c.ptr, out, (int64_t)n, (int64_t)n, "mmult");
}

int pyxrd_leaf_matrix_powers(pyxrd_ctx *ctx, const double *q, int64_t n_points, int rank, int n_max, double *out) {
    if (!ctx || !q || !out || n_points < 0 || rank <= 0 || n_max < 0) return 1;
    CK(cudaSetDevice(ctx->device));
    const size_t per = (size_t)n_points * rank * rank * 2, n = per * ((size_t)n_max + 1);
    if (per == 0) return 0;
    if (upload(ctx, ctx->d_leaf_a, q, per)) return 1;
    CK(ctx->d_leaf_c.ensure(sizeof(double) * n));
    const long long threads = (long long)n_points * rank;
    k_leaf_matrix_powers<<<(unsigned)((threads + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_leaf_a.as<double2>(), (long long)n_points, rank,
                                                                                   n_max, ctx->d_leaf_c.as<double2>());
    LAUNCH_CHECK("k_leaf_matrix_powers");
    return read_back(ctx, ctx->d_leaf_c.ptr, out, (int64_t)n, (int64_t)n, "matrix powers");
}

// ---- population from a template + binding table (host side: layout only) ------------------------
}  // extern "C"

struct pyxrd_expanded {
    pyxrd_batch_desc desc{};
    std::vector<int32_t> phase_i, phase_comp, comp_i, atom_type, job_phase, prob_model;
    std::vector<double> phase_d, matrices, comp_d, atom_d, fractions, scales, bgshifts, prob_params;
};

namespace {
thread_local std::string g_expand_error;

template <class T>
void replicate(std::vector<T> &dst, const T *src, size_t n, int K) {
    dst.resize(n * (size_t)K);
    for (int k = 0; k < K; ++k)
        if (n) memcpy(dst.data() + (size_t)k * n, src, n * sizeof(T));
}
}  // namespace

extern "C" {

const char *pyxrd_expand_error(void) { return g_expand_error.c_str(); }

int pyxrd_expand_population(const pyxrd_batch_desc *t, const pyxrd_population_desc *pop, pyxrd_expanded **out) {
    auto bad = [&](const std::string &msg) { g_expand_error = msg; return 1; };
    if (!t || !pop || !out) return bad("expand_population: null argument");
    *out = nullptr;
    if (t->n_candidates != 1) return bad("expand_population: the template must hold exactly one candidate");
    const int K = pop->n_candidates, D = pop->n_params, S = pop->n_specimens, Z = pop->n_patterns;
    if (K <= 0 || D < 0 || pop->n_bindings < 0 || S <= 0 || Z <= 0) return bad("expand_population: need K, S, Z > 0");
    const int NP = t->n_phases, NC = t->n_components, NA = t->n_atoms, NPC = t->n_phase_comp, M = t->n_phase_slots;
    const int64_t NM = t->n_matrix;
    const size_t JPC = (size_t)S * Z * M;
    if ((int64_t)NP * K > 0x7fffffffLL || (int64_t)NA * K > 0x7fffffffLL || NM * K > 0x7fffffffLL)
        return bad("expand_population: population too large for 32-bit indices");
    pyxrd_expanded *e = new pyxrd_expanded();
    replicate(e->phase_i, t->phase_i, (size_t)NP * PYXRD_PHASE_I_STRIDE, K);
    replicate(e->phase_d, t->phase_d, (size_t)NP * PYXRD_PHASE_D_STRIDE, K);
    replicate(e->phase_comp, t->phase_comp, (size_t)NPC, K);
    replicate(e->matrices, t->matrices, (size_t)NM, K);
    replicate(e->comp_d, t->comp_d, (size_t)NC * PYXRD_COMP_D_STRIDE, K);
    replicate(e->comp_i, t->comp_i, (size_t)NC * PYXRD_COMP_I_STRIDE, K);
    replicate(e->atom_d, t->atom_d, (size_t)NA * PYXRD_ATOM_D_STRIDE, K);
    replicate(e->atom_type, t->atom_type, (size_t)NA, K);
    replicate(e->job_phase, t->job_phase, JPC, K);
    replicate(e->fractions, t->fractions, (size_t)M, K);
    replicate(e->scales, t->scales, (size_t)S, K);
    replicate(e->bgshifts, t->bgshifts, (size_t)S, K);
    const bool model_driven = t->prob_model && t->prob_params;
    if (model_driven) {
        replicate(e->prob_model, t->prob_model, (size_t)NP, K);
        replicate(e->prob_params, t->prob_params, (size_t)NP * PYXRD_PROB_STRIDE, K);
    }
    // the indices that point into per-candidate tables move with the candidate
    for (int k = 1; k < K; ++k) {
        for (int p = 0; p < NP; ++p) {
            int32_t *pi = e->phase_i.data() + ((size_t)k * NP + p) * PYXRD_PHASE_I_STRIDE;
            pi[5] += k * NPC;
            pi[6] += (int32_t)(k * NM);
        }
        for (int i = 0; i < NPC; ++i) e->phase_comp[(size_t)k * NPC + i] += k * NC;
        for (int c = 0; c < NC; ++c) e->comp_i[((size_t)k * NC + c) * PYXRD_COMP_I_STRIDE] += k * NA;
        for (size_t j = 0; j < JPC; ++j)
            if (e->job_phase[(size_t)k * JPC + j] >= 0) e->job_phase[(size_t)k * JPC + j] += k * NP;
    }
    // slot = scale * x[param] + offset
    struct Target { double *base; size_t per_candidate; };
    const Target targets[8] = {{e->phase_d.data(), (size_t)NP * PYXRD_PHASE_D_STRIDE}, {e->comp_d.data(), (size_t)NC * PYXRD_COMP_D_STRIDE},
                               {e->atom_d.data(), (size_t)NA * PYXRD_ATOM_D_STRIDE}, {e->matrices.data(), (size_t)NM},
                               {e->prob_params.data(), model_driven ? (size_t)NP * PYXRD_PROB_STRIDE : 0}, {e->fractions.data(), (size_t)M},
                               {e->scales.data(), (size_t)S}, {e->bgshifts.data(), (size_t)S}};
    for (int b = 0; b < pop->n_bindings; ++b) {
        const pyxrd_binding &bd = pop->bindings[b];
        if (bd.target < 0 || bd.target > 7 || bd.param < 0 || bd.param >= D || bd.index < 0 || bd.mode < 0 || bd.mode > 2 ||
            (size_t)bd.index >= targets[bd.target].per_candidate) {
            delete e;
            return bad("expand_population: binding " + std::to_string(b) + " is out of range");
        }
        const Target &tg = targets[bd.target];
        for (int k = 0; k < K; ++k) {
            double &slot = tg.base[(size_t)k * tg.per_candidate + bd.index];
            const double v = bd.scale * pop->x[(size_t)k * D + bd.param] + bd.offset;
            slot = bd.mode == PYXRD_BIND_MODE_MULTIPLY ? slot * v : bd.mode == PYXRD_BIND_MODE_ADD ? slot + v : v;
        }
    }
    if (pop->csds_auto_maximum)
        for (int k = 0; k < K; ++k)
            for (int p = 0; p < NP; ++p)
                if (pop->csds_auto_maximum[p])      // maximum = int(factor * average)  (phases/models/CSDS.py:127,140)
                    e->phase_i[((size_t)k * NP + p) * PYXRD_PHASE_I_STRIDE + 3] =
                        (int32_t)(pop->csds_factor * e->phase_d[((size_t)k * NP + p) * PYXRD_PHASE_D_STRIDE + 1]);
    pyxrd_batch_desc &d = e->desc;
    d.n_candidates = K; d.n_phase_slots = M; d.n_phases = NP * K; d.n_components = NC * K; d.n_atoms = NA * K;
    d.n_phase_comp = NPC * K; d.n_matrix = NM * K;
    d.phase_i = e->phase_i.data(); d.phase_d = e->phase_d.data(); d.phase_comp = e->phase_comp.data();
    d.matrices = e->matrices.data(); d.comp_d = e->comp_d.data(); d.comp_i = e->comp_i.data();
    d.atom_d = e->atom_d.data(); d.atom_type = e->atom_type.data(); d.job_phase = e->job_phase.data();
    d.fractions = e->fractions.data(); d.scales = e->scales.data(); d.bgshifts = e->bgshifts.data();
    d.prob_model = model_driven ? e->prob_model.data() : nullptr;
    d.prob_params = model_driven ? e->prob_params.data() : nullptr;
    *out = e;
    return 0;
}

void pyxrd_expanded_desc(const pyxrd_expanded *e, pyxrd_batch_desc *out) {
    if (e && out) *out = e->desc;
}

void pyxrd_expanded_free(pyxrd_expanded *e) { delete e; }

}  // extern "C"

namespace {
// The automatic CSDS maximum int(factor * average) depends on the bound average: its range over the solution vectors
// x[K][D] sizes the coefficient table (*ncap) and the shared memory of each job class (the last binding of a slot wins,
// as in the expansion).
int population_caps(pyxrd_ctx *ctx, const double *x, int K, int D, int *ncap) {
    pyxrd_ctx::Population &pp = ctx->pop;
    const int NP = (int)(pp.t_phase_i.size() / PYXRD_PHASE_I_STRIDE), NB = (int)pp.bind.size();
    std::vector<int> pmax(NP);
    for (int p = 0; p < NP; ++p) pmax[p] = pp.t_phase_i[(size_t)p * PYXRD_PHASE_I_STRIDE + 3];
    if (!pp.auto_max.empty())
        for (int p = 0; p < NP; ++p) {
            if (!pp.auto_max[p]) continue;
            const pyxrd_binding *last = nullptr;
            for (int b = 0; b < NB; ++b)
                if (pp.bind[b].target == PYXRD_BIND_PHASE_D && pp.bind[b].index == p * PYXRD_PHASE_D_STRIDE + 1) last = &pp.bind[b];
            int lo, hi;
            if (!last) {
                lo = hi = (int32_t)(pp.factor * pp.t_phase_d[(size_t)p * PYXRD_PHASE_D_STRIDE + 1]);
            } else {
                lo = 0x7fffffff; hi = -0x7fffffff;
                for (int k = 0; k < K; ++k) {
                    const double avg = last->scale * x[(size_t)k * D + last->param] + last->offset;
                    const double mx = pp.factor * avg;
                    if (!(mx > -1e9 && mx < 1e9)) return fail(ctx, "set_population: candidate %d: CSDS average %g of phase %d is not usable", k, avg, p);
                    lo = std::min(lo, (int)mx);
                    hi = std::max(hi, (int)mx);
                }
            }
            pmax[p] = hi;
            const int32_t *pi = pp.t_phase_i.data() + (size_t)p * PYXRD_PHASE_I_STRIDE;
            if ((pi[4] & PYXRD_FLAG_VALID_PROBS) && !(pi[4] & PYXRD_FLAG_RAW_PATTERN) && lo < pi[2])
                return fail(ctx, "set_batch: phase %d CSDS range [%d, %d] unsupported", p, pi[2], lo);
        }
    int cap = pp.ncap0;
    for (size_t c = 0; c < ctx->classes.size(); ++c) {
        RankClass &rc = ctx->classes[c];
        rc.max_ntop_cap = pp.cls_cap0[c];
        for (int j : pp.cls_jobs[c]) {
            const int pb = pp.t_job_pb[j];
            const int32_t *pi = pp.t_phase_i.data() + (size_t)pb * PYXRD_PHASE_I_STRIDE;
            if (!(pi[4] & PYXRD_FLAG_RAW_PATTERN)) {
                rc.max_ntop_cap = std::max(rc.max_ntop_cap, pmax[pb] + 1);
                cap = std::max(cap, pmax[pb] + 2);
            }
        }
    }
    *ncap = cap;
    return 0;
}
}  // namespace

extern "C" {

int pyxrd_set_population(pyxrd_ctx *ctx, const pyxrd_batch_desc *t, const pyxrd_population_desc *pop) {
    if (!ctx) return 1;
    if (!t || !pop) return fail(ctx, "set_population: null argument");
    if (!ctx->have_static) return fail(ctx, "set_population: call pyxrd_set_static first");
    const StaticDev &st = ctx->st;
    if (pop->n_specimens != st.S || pop->n_patterns != st.Z)
        return fail(ctx, "set_population: S / Z of the population (%d, %d) differ from the static tables (%d, %d)",
                    pop->n_specimens, pop->n_patterns, st.S, st.Z);
    if (t->n_candidates != 1) return fail(ctx, "set_population: the template must hold exactly one candidate");
    const int K = pop->n_candidates, D = pop->n_params, NB = pop->n_bindings, S = st.S, Z = st.Z;
    if (K <= 0 || D < 0 || NB < 0 || t->n_phase_slots <= 0) return fail(ctx, "set_population: need K > 0, M > 0");
    if (NB > 0 && (!pop->bindings || !pop->x)) return fail(ctx, "set_population: bindings without a table / solution vectors");
    CK(cudaSetDevice(ctx->device));
    ctx->have_batch = false;
    ctx->have_intensities = false;
    ctx->have_opt = false;
    ctx->have_totals = false;
    const int NP = t->n_phases, NC = t->n_components, NA = t->n_atoms, NPC = t->n_phase_comp, M = t->n_phase_slots;
    const int64_t NM = t->n_matrix;
    if ((int64_t)NP * K > 0x7fffffffLL || (int64_t)NA * K > 0x7fffffffLL || NM * K > 0x7fffffffLL)
        return fail(ctx, "set_population: population too large for 32-bit indices");
    int ncap = 2, acap = 1;
    if (validate_batch(ctx, t, &ncap, &acap)) return 1;
    const bool model_driven = t->prob_model && t->prob_params;
    const size_t per[8] = {(size_t)NP * PYXRD_PHASE_D_STRIDE, (size_t)NC * PYXRD_COMP_D_STRIDE, (size_t)NA * PYXRD_ATOM_D_STRIDE, (size_t)NM,
                           model_driven ? (size_t)NP * PYXRD_PROB_STRIDE : 0, (size_t)M, (size_t)S, (size_t)S};
    for (int b = 0; b < NB; ++b) {
        const pyxrd_binding &bd = pop->bindings[b];
        if (bd.target < 0 || bd.target > 7 || bd.param < 0 || bd.param >= D || bd.index < 0 || bd.mode < 0 || bd.mode > 2 ||
            (size_t)bd.index >= per[bd.target])
            return fail(ctx, "set_population: binding %d is out of range", b);
        if (bd.mode != PYXRD_BIND_MODE_SET && pop->csds_auto_maximum && bd.target == PYXRD_BIND_PHASE_D &&
            bd.index % PYXRD_PHASE_D_STRIDE == 1 && pop->csds_auto_maximum[bd.index / PYXRD_PHASE_D_STRIDE])
            return fail(ctx, "set_population: binding %d: a CSDS average with an automatic maximum takes one refinable", b);
    }
    // template job tables and classes (K = 1); every candidate repeats them
    std::vector<int> job_pb, job_spec, jobids;
    std::vector<long long> job_out;
    if (build_jobs(ctx, t, job_pb, job_spec, job_out, jobids)) return 1;
    // remember the template side of the population: pyxrd_update_population re-expands it for new solution vectors
    pyxrd_ctx::Population &pp = ctx->pop;
    pp.valid = false;
    pp.t_phase_i.assign(t->phase_i, t->phase_i + (size_t)NP * PYXRD_PHASE_I_STRIDE);
    pp.t_phase_d.assign(t->phase_d, t->phase_d + (size_t)NP * PYXRD_PHASE_D_STRIDE);
    pp.t_job_pb = job_pb;
    pp.bind.assign(pop->bindings, pop->bindings + NB);
    pp.auto_max.clear();
    if (pop->csds_auto_maximum) pp.auto_max.assign(pop->csds_auto_maximum, pop->csds_auto_maximum + NP);
    pp.factor = pop->csds_factor;
    pp.ncap0 = ncap;
    pp.acap = acap;
    pp.cls_jobs.clear();
    pp.cls_cap0.clear();
    const size_t JPC = job_pb.size();
    std::vector<int> cls_first, cls_n;
    {
        size_t pos = 0;
        for (RankClass &rc : ctx->classes) {
            cls_first.push_back((int)pos);
            cls_n.push_back((int)rc.jobs.size());
            pos += rc.jobs.size();
            pp.cls_jobs.push_back(rc.jobs);
            pp.cls_cap0.push_back(rc.max_ntop_cap);
            rc.per_cand = rc.jobs.size();
            rc.uniform = true;
            rc.jobs.clear();
        }
    }
    if (population_caps(ctx, pop->x, K, D, &ncap)) return 1;
    const int NCLS = (int)ctx->classes.size();
    // layer classes of the template's components; a component whose layer atoms a refinable drives keeps its layer
    // inside the phase kernels.  Every template component stands for K components of the population.
    std::vector<char> bound_comp((size_t)std::max(NC, 1), 0);
    for (int b = 0; b < NB; ++b) {
        const pyxrd_binding &bd = pop->bindings[b];
        if (bd.target != PYXRD_BIND_ATOM_D) continue;
        const int atom = bd.index / PYXRD_ATOM_D_STRIDE;
        for (int c = 0; c < NC; ++c) {
            const int32_t *ci = t->comp_i + (size_t)c * PYXRD_COMP_I_STRIDE;
            if (atom >= ci[0] && atom < ci[0] + ci[1]) bound_comp[c] = 1;
        }
    }
    std::vector<int32_t> comp_class, class_rep;
    const int n_layer_classes = layer_classes(t, &bound_comp, ctx->use_layer_tables ? K : 0, comp_class, class_rep);
    class_rep.resize(PX_MAX_LAYER_CLASSES, 0);

    BatchSizes z;
    z.n_classes = n_layer_classes;
    z.K = K; z.M = M; z.NP = NP * K; z.NC = NC * K; z.NA = NA * K; z.NPC = NPC * K; z.NM = NM * K;
    z.NJ = (long long)JPC * K; z.njobids = jobids.size() * (size_t)K; z.model_driven = model_driven;
    if (z.NJ > 0x7fffffffLL) return fail(ctx, "set_batch: too many patterns (%lld)", z.NJ);
    const BlobLayout L = plan_blob(z, S);
    CK(ctx->d_blob.ensure(L.total));

    // ---- template + solution vectors + table: one pinned staging blob, one H2D copy ------------------
    struct Part { const void *src; size_t bytes, off; };
    std::vector<Part> parts;
    size_t total = 0;
    auto add = [&](const void *src, size_t bytes) {
        parts.push_back(Part{src, src ? bytes : 0, total});
        total = align_up(total + std::max<size_t>(bytes, 8), 256);
        return (int)parts.size() - 1;
    };
    const int a_phase_i = add(t->phase_i, sizeof(int32_t) * (size_t)NP * PYXRD_PHASE_I_STRIDE);
    const int a_phase_d = add(t->phase_d, sizeof(double) * (size_t)NP * PYXRD_PHASE_D_STRIDE);
    const int a_phase_comp = add(t->phase_comp, sizeof(int32_t) * (size_t)NPC);
    const int a_matrices = add(t->matrices, sizeof(double) * (size_t)NM);
    const int a_comp_d = add(t->comp_d, sizeof(double) * (size_t)NC * PYXRD_COMP_D_STRIDE);
    const int a_comp_i = add(t->comp_i, sizeof(int32_t) * (size_t)NC * PYXRD_COMP_I_STRIDE);
    const int a_atom_d = add(t->atom_d, sizeof(double) * (size_t)NA * PYXRD_ATOM_D_STRIDE);
    const int a_atom_t = add(t->atom_type, sizeof(int32_t) * (size_t)NA);
    const int a_job_pb = add(job_pb.data(), sizeof(int) * JPC);
    const int a_job_spec = add(job_spec.data(), sizeof(int) * JPC);
    const int a_job_out = add(job_out.data(), sizeof(long long) * JPC);
    const int a_jobids = add(jobids.data(), sizeof(int) * jobids.size());
    const int a_frac = add(t->fractions, sizeof(double) * (size_t)M);
    const int a_scales = add(t->scales, sizeof(double) * (size_t)S);
    const int a_bg = add(t->bgshifts, sizeof(double) * (size_t)S);
    const int a_pmodel = add(model_driven ? t->prob_model : nullptr, sizeof(int32_t) * (size_t)NP);
    const int a_pparam = add(model_driven ? t->prob_params : nullptr, sizeof(double) * (size_t)NP * PYXRD_PROB_STRIDE);
    const int a_comp_class = add(comp_class.data(), sizeof(int32_t) * (size_t)NC);
    const int a_class_rep = add(class_rep.data(), sizeof(int32_t) * PX_MAX_LAYER_CLASSES);
    const int a_cls_first = add(cls_first.data(), sizeof(int) * (size_t)NCLS);
    const int a_cls_n = add(cls_n.data(), sizeof(int) * (size_t)NCLS);
    const int a_x = add(pop->x, sizeof(double) * (size_t)K * D);
    const int a_bind = add(pop->bindings, sizeof(pyxrd_binding) * (size_t)NB);
    const int a_auto = add(pop->csds_auto_maximum, sizeof(uint8_t) * (size_t)NP);
    CK(cudaStreamSynchronize(ctx->stream));        // the staging blob may still feed a previous copy
    CK(ctx->h_tmpl.ensure(total));
    CK(ctx->d_tmpl.ensure(total));
    for (const Part &p : parts)
        if (p.bytes) memcpy(static_cast<char *>(ctx->h_tmpl.ptr) + p.off, p.src, p.bytes);
    CK(cudaMemcpyAsync(ctx->d_tmpl.ptr, ctx->h_tmpl.ptr, total, cudaMemcpyHostToDevice, ctx->stream));

    ExpandDev e{};
    e.K = K; e.D = D; e.NB = NB; e.NP = NP; e.NC = NC; e.NA = NA; e.NPC = NPC; e.M = M; e.S = S; e.JPC = (int)JPC; e.NCLS = NCLS;
    e.NM = NM; e.cand_stride = (long long)M * Z * st.sumX;
    auto tp = [&](int idx) { return static_cast<char *>(ctx->d_tmpl.ptr) + parts[idx].off; };
    auto dp = [&](int idx) { return static_cast<char *>(ctx->d_blob.ptr) + L.off[idx]; };
    e.t_phase_i = (const int *)tp(a_phase_i); e.t_phase_d = (const double *)tp(a_phase_d); e.t_phase_comp = (const int *)tp(a_phase_comp);
    e.t_matrices = (const double *)tp(a_matrices); e.t_comp_d = (const double *)tp(a_comp_d); e.t_comp_i = (const int *)tp(a_comp_i);
    e.t_atom_d = (const double *)tp(a_atom_d); e.t_atom_t = (const int *)tp(a_atom_t); e.t_job_pb = (const int *)tp(a_job_pb);
    e.t_job_spec = (const int *)tp(a_job_spec); e.t_job_out = (const long long *)tp(a_job_out); e.t_jobids = (const int *)tp(a_jobids);
    e.t_frac = (const double *)tp(a_frac); e.t_scales = (const double *)tp(a_scales); e.t_bg = (const double *)tp(a_bg);
    e.t_pmodel = (const int *)tp(a_pmodel); e.t_pparam = (const double *)tp(a_pparam);
    e.t_comp_class = (const int *)tp(a_comp_class);
    e.comp_class = (int *)dp(P_COMP_CLASS);
    e.cls_first = (const int *)tp(a_cls_first); e.cls_n = (const int *)tp(a_cls_n);
    e.x = (const double *)tp(a_x); e.bind = (const pyxrd_binding *)tp(a_bind);
    e.auto_max = pop->csds_auto_maximum ? (const unsigned char *)tp(a_auto) : nullptr;
    e.factor = pop->csds_factor;
    e.phase_i = (int *)dp(P_PHASE_I); e.phase_d = (double *)dp(P_PHASE_D); e.phase_comp = (int *)dp(P_PHASE_COMP);
    e.matrices = (double *)dp(P_MATRICES); e.comp_d = (double *)dp(P_COMP_D); e.comp_i = (int *)dp(P_COMP_I);
    e.atom_d = (double *)dp(P_ATOM_D); e.atom_t = (int *)dp(P_ATOM_T); e.job_pb = (int *)dp(P_JOB_PB); e.job_spec = (int *)dp(P_JOB_SPEC);
    e.job_out = (long long *)dp(P_JOB_OUT); e.jobids = (int *)dp(P_JOBIDS); e.frac = (double *)dp(P_FRAC); e.scales = (double *)dp(P_SCALES);
    e.bg = (double *)dp(P_BG); e.pmodel = model_driven ? (int *)dp(P_PMODEL) : nullptr; e.pparam = model_driven ? (double *)dp(P_PPARAM) : nullptr;
    // the class representatives are components of candidate 0 (= the template's own indices)
    CK(cudaMemcpyAsync(dp(P_CLASS_REP), tp(a_class_rep), sizeof(int32_t) * PX_MAX_LAYER_CLASSES, cudaMemcpyDeviceToDevice, ctx->stream));
    k_expand_population<<<K, 128, 0, ctx->stream>>>(e);
    LAUNCH_CHECK("k_expand_population");
    if (bind_batch(ctx, z, L, ncap, acap)) return 1;
    pp.n_layer_classes = n_layer_classes;
    pp.e = e;
    pp.x_off = parts[a_x].off;
    pp.NJ = z.NJ;
    pp.njobids = z.njobids;
    pp.M = M;
    pp.model_driven = model_driven;
    pp.valid = true;
    return 0;
}

int pyxrd_update_population(pyxrd_ctx *ctx, const double *x) {
    if (!ctx) return 1;
    pyxrd_ctx::Population &pp = ctx->pop;
    if (!pp.valid || !ctx->have_static) return fail(ctx, "update_population: no resident population (call pyxrd_set_population first)");
    CK(cudaSetDevice(ctx->device));
    const ExpandDev &e = pp.e;
    const int K = e.K, D = e.D;
    ctx->have_batch = false;
    ctx->have_intensities = false;
    ctx->have_opt = false;
    ctx->have_totals = false;
    int ncap = pp.ncap0;
    if (x) {
        pp.valid = false;                              // stays invalid if anything below fails
        if (population_caps(ctx, x, K, D, &ncap)) return 1;
        const size_t bytes = sizeof(double) * (size_t)K * D;
        CK(cudaStreamSynchronize(ctx->stream));        // the staging blob may still feed the previous copy
        char *stage = static_cast<char *>(ctx->h_tmpl.ptr) + pp.x_off;
        memcpy(stage, x, bytes);
        CK(cudaMemcpyAsync(static_cast<char *>(ctx->d_tmpl.ptr) + pp.x_off, stage, bytes, cudaMemcpyHostToDevice, ctx->stream));
    } else {
        // same solution vectors: the caps of the resident population stand
        ncap = ctx->bt.ncap;
    }
    BatchSizes z;
    z.K = K; z.M = pp.M; z.NP = e.NP * K; z.NC = e.NC * K; z.NA = e.NA * K; z.NPC = e.NPC * K; z.NM = e.NM * K;
    z.NJ = pp.NJ; z.njobids = pp.njobids; z.model_driven = pp.model_driven; z.n_classes = pp.n_layer_classes;
    const BlobLayout L = plan_blob(z, ctx->st.S);
    k_expand_population<<<K, 128, 0, ctx->stream>>>(e);
    LAUNCH_CHECK("k_expand_population");
    // the layer tables depend on the template's unbound layer rows and the static tables only: still valid
    if (bind_batch(ctx, z, L, ncap, pp.acap, true)) return 1;
    pp.valid = true;
    return 0;
}

int pyxrd_evaluate_population(pyxrd_ctx *ctx, const pyxrd_batch_desc *t, const pyxrd_population_desc *pop, double *residuals_out) {
    if (pyxrd_set_population(ctx, t, pop)) return 1;
    if (pyxrd_compute_all(ctx, PYXRD_RESIDUALS_ONLY)) return 1;
    return pyxrd_read_residuals(ctx, residuals_out, (int64_t)ctx->bt.K * (ctx->st.S + 1));
}

int pyxrd_optimize_population(pyxrd_ctx *ctx, const pyxrd_batch_desc *t, const pyxrd_population_desc *pop, int refine_bg,
                              int outer_iterations, int newton_iterations, double *residuals_out, double *solutions_out) {
    if (pyxrd_set_population(ctx, t, pop)) return 1;
    if (pyxrd_compute_intensities(ctx)) return 1;
    if (pyxrd_optimize(ctx, refine_bg, outer_iterations, newton_iterations)) return 1;
    if (pyxrd_read_opt_residuals(ctx, residuals_out, (int64_t)ctx->bt.K * (ctx->st.S + 1))) return 1;
    if (solutions_out) return pyxrd_read_solutions(ctx, solutions_out, (int64_t)ctx->bt.K * (ctx->bt.M + 2 * ctx->st.S));
    return 0;
}

int pyxrd_leaf_probabilities(pyxrd_ctx *ctx, int model, int G, const double *params, double *W_out, double *P_out,
                             int *valid_out, int *rank_out) {
    if (!ctx || !params) return 1;
    const int want = pyxrd_prob_rank(model, G);
    if (want < 1)
        return fail(ctx, "leaf_probabilities: model %d with G = %d is not available", model, G);
    CK(cudaSetDevice(ctx->device));
    if (upload(ctx, ctx->d_leaf_a, params, PYXRD_PROB_STRIDE)) return 1;
    CK(ctx->d_leaf_b.ensure(sizeof(double) * 2 * PX_PROB_MAXSQ));
    CK(ctx->d_leaf_c.ensure(sizeof(int) * 2));
    k_leaf_probabilities<<<1, 32, 0, ctx->stream>>>(model, G, ctx->d_leaf_a.as<double>(), ctx->d_leaf_b.as<double>(),
                                                     ctx->d_leaf_b.as<double>() + PX_PROB_MAXSQ, ctx->d_leaf_c.as<int>());
    LAUNCH_CHECK("k_leaf_probabilities");
    int flags[2] = {0, 0};
    CK(cudaMemcpyAsync(flags, ctx->d_leaf_c.ptr, sizeof(flags), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(W_out, ctx->d_leaf_b.ptr, sizeof(double) * want * want, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(P_out, ctx->d_leaf_b.as<double>() + PX_PROB_MAXSQ, sizeof(double) * want * want, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (valid_out) *valid_out = flags[0];
    if (rank_out) *rank_out = flags[1];
    return 0;
}

int pyxrd_stage_times(pyxrd_ctx *ctx, double *ms_out) {
    if (!ctx || !ms_out) return 1;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    const int pairs[4][2] = {{0, 1}, {2, 3}, {4, 5}, {6, 7}};
    for (int i = 0; i < 4; ++i) {
        float ms = 0.f;
        if (ctx->ev_set[i]) CK(cudaEventElapsedTime(&ms, ctx->ev[pairs[i][0]], ctx->ev[pairs[i][1]]));
        ms_out[i] = ms;
    }
    return 0;
}

int pyxrd_dfma_peak(pyxrd_ctx *ctx, int iterations, double *tflops_out, double *ms_out) {
    if (!ctx) return 1;
    CK(cudaSetDevice(ctx->device));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    const int blocks = sms * 8, threads = 256;
    CK(ctx->d_leaf_c.ensure(sizeof(double) * (size_t)blocks * threads));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    k_dfma_peak<<<blocks, threads, 0, ctx->stream>>>(ctx->d_leaf_c.as<double>(), 16, 1.0);   // warm-up
    LAUNCH_CHECK("k_dfma_peak");
    CK(cudaEventRecord(e0, ctx->stream));
    k_dfma_peak<<<blocks, threads, 0, ctx->stream>>>(ctx->d_leaf_c.as<double>(), iterations, 1.0);
    LAUNCH_CHECK("k_dfma_peak");
    CK(cudaEventRecord(e1, ctx->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double flops = 2.0 * 8.0 * 16.0 * (double)iterations * (double)blocks * threads;
    if (tflops_out) *tflops_out = flops / (ms * 1e-3) * 1e-12;
    if (ms_out) *ms_out = ms;
    return 0;
}

}  // extern "C"
