// Device code of the pyxrd_b200 hot path (sm_100a, fp64).
//
// Kernel map (reference = pyxrd/calculations/*, v0.8.4):
//   k_specimen_tables   specimen.py:21-43,50  goniometer.py:36-37      per-point static tables
//   k_asf_table         atoms.py:13-38                                   ASF per (atom type, point)
//   k_phase_prologue    CSDS.py:13-46 math_tools.py:36-37 phases.py:48-66,147-151
//   k_component_z       components.py:15-16,40-50                        interlayer z-stretch
//   k_phase_small<R,G>  atoms.py:40-67 components.py:18-54 phases.py:20-39,106-158
//   k_phase_staged<R,G> goniometer.py:15-34 phases.py:81-95 specimen.py:45-62   (ranks 16, 27)
//   k_phase_generic     same, any rank <= 64
//   k_wavelength        specimen.py:64-74
//   k_mixture_residual  specimen.py:84-88,16-19 statistics.py:22-27 mixture.py:242-253
//
// The series: with u_I = SF_comp(I), F = conj(u) u^T, W, S = mean*Id + sum_n coef_n Q^n,
// Q = P diag(phi), the reference's Re Tr(F W S) equals Re( a^T (S b) ) with a = W^T u,
// b = conj(u); S b is evaluated by the vector Horner recurrence y <- P (phi o (y + coef_n b)),
// O(N r^2) real x complex MACs per point instead of the reference's O(N r^3) matrix powers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <type_traits>

#define PX_TPB 128          // threads (= 2-theta points) per block in the phase kernels
#define PX_GEN_TPB 64       // generic-rank fallback
#define PX_TABLE_PAD 1024    // doubles of slack after every per-point table (multi-point threads read past a pattern end)
#define PX_IMG_HDR 74        // 8-byte words of the header block of PhaseSmem = first words of a phase image
#define PX_TWO_PI 6.283185307179586        // python: 2 * math.pi
#define PX_PI 3.141592653589793
#define PX_SQRT2PI 2.5066282746310002      // math.sqrt(2 * pi)
#define PX_SQRT8 2.8284271247461903        // math.sqrt(8)

// (constants of the phase kernels live in constant memory so that ptxas keeps them in uniform registers across the
// atom loop instead of re-materialising every 64-bit immediate with two UMOVs: 25 % of all issued instructions of the
// first version of k_phase_small)

// erf(Q) and exp(-Q^2) of the Lorentz-polarisation factor as piecewise degree-7 polynomials (tools/make_lpf_table.py):
// [i / 16, (i + 1) / 16) for i < 104, one 128-byte line of 8 (erf, exp) coefficient pairs per interval, tail entry = (1, 0)
#define PX_LPF_TABLE_DECL __device__ __align__(128) const double PX_LPF_TAB
#include "lpf_table.inc"

// table-assisted sine / cosine of k_phase_small (tools/make_sincos_table.py): reduction constants PX_ST, table PX_SCT
#define PX_SCT_N 128
#ifndef PX_SCT_SMEM
#define PX_SCT_SMEM true       // k_phase_small reads its own copy of the table in shared memory (false: the global one)
#endif
#include "sincos_table.inc"

// exp(a), a <= 0: a = n ln 2 + r, |r| <= ln 2 / 2; Taylor polynomial of degree 13 (remainder 4e-18), 2^n into the exponent
__constant__ double PX_EX[18] = {
    1.4426950408889634,         // log2(e)
    6755399441055744.0,         // 1.5 * 2^52
    -6.93147180369123816490e-01, -1.90821492927058770002e-10,      // -ln 2 in two parts (fdlibm)
    1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0, 1.0 / 40320.0, 1.0 / 5040.0,
    1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, 1.0, 1.0};

template <int N>
__device__ __forceinline__ void px_exp_neg_n(const double (&a)[N], double (&out)[N]) {
    double r[N], p[N];
    int n[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double ac = fmax(a[k], -700.0);                  // the result stays a normal number
        const double t = fma(ac, PX_EX[0], PX_EX[1]);
        n[k] = __double2loint(t);
        const double nf = t - PX_EX[1];
        r[k] = fma(nf, PX_EX[2], ac);
        r[k] = fma(nf, PX_EX[3], r[k]);
        p[k] = PX_EX[4];
    }
#pragma unroll
    for (int j = 5; j <= 17; ++j)
#pragma unroll
        for (int k = 0; k < N; ++k) p[k] = fma(p[k], r[k], PX_EX[j]);
#pragma unroll
    for (int k = 0; k < N; ++k) out[k] = __hiloint2double(__double2hiint(p[k]) + (n[k] << 20), __double2loint(p[k]));
}

// sin and cos of a (|a| < 2^20): two-part Cody-Waite reduction with FMAs to |r| <= pi / 128 around a table entry
__device__ __forceinline__ void px_sincos(double a, double &s, double &c) {
    // table-assisted (see px_sincos_n): a = k h + r, h = pi / 64, (S, C) = (sin, cos)(k h) from PX_SCT
    const double t = fma(a, PX_ST[0], PX_ST[1]);
    const double q = t - PX_ST[1];
    const double2 sc = __ldg(reinterpret_cast<const double2 *>(PX_SCT) + (__double2loint(t) & (PX_SCT_N - 1)));
    double r = fma(q, PX_ST[2], a);
    r = fma(q, PX_ST[3], r);
    const double r2 = r * r;
    const double ps = fma(fma(r2, PX_ST[4], PX_ST[5]), r2, PX_ST[6]), pc = fma(fma(r2, PX_ST[7], PX_ST[8]), r2, PX_ST[9]);
    const double sr = fma(r * r2, ps, r), cm1 = r2 * pc;
    s = fma(sc.y, sr, fma(sc.x, cm1, sc.x));
    c = fma(-sc.x, sr, fma(sc.y, cm1, sc.y));
}

struct StaticDev {
    int S, T, Z, maxX;
    long long sumX;
    const int *spec_npts;        // [S]
    const long long *spec_off;   // [S+1]
    const double *gonio;         // [S][12]
    const double *derived;       // [S][8]: pol, S_soller, S_soller^2, obs_den
    const double *theta, *sin_t, *inv_sin, *stl, *c2t2, *corr;   // [sumX]
    const double *asf;           // [T][sumX]
    const int *spec_nwl;         // [S]
    const int *wl_first;         // [S] first entry in wl_fraction
    const long long *wl_off;     // [S] offset of (s, j=0, x=0) in wl_index / weights
    const double *wl_fraction;
    const int *wl_index;
    const double *wl_w_hi, *wl_w_lo;
    const double *observed;      // [S][Z][X_s] at Z*spec_off[s]
    const unsigned char *selected;
    const double *atypes;        // [T][12]
};

struct BatchDev {
    int K, M, NP, NC, NA, ncap;      // ncap = stride of coef
    long long cand_stride;           // doubles per candidate in the intensity buffer = M*Z*sumX
    const int *phase_i;              // [NP][8]
    const double *phase_d;           // [NP][8]
    const int *phase_comp;
    const double *matrices;
    const double *comp_d;            // [NC][8]
    const int *comp_i;               // [NC][4]
    const double *atom_d;            // [NA][2]
    const int *atom_type;            // [NA]
    double *atom_z;                  // [NA]   derived: stretched z
    int *atom_plane;                 // [NA]   derived: plane index within the component (runs of equal z)
    int *comp_nplanes;               // [NC]   derived
    int *comp_ilplane;               // [NC]   derived: plane index of the first interlayer atom (= comp_nplanes without any)
    // candidate-invariant layer stacks: the silicate layer of a component (atoms, contents, positions: calculate_z only
    // moves INTERLAYER atoms, components.py:15-16,40-45) is the same for every candidate unless a refinable drives it;
    // components with identical layer rows form a class whose part of the structure factor,
    // sum_layer pn asf(s) exp(2 pi i z s), is tabulated once per batch per point (k_layer_table)
    const int *comp_class;           // [NC]   class of the component's layer, -1: evaluated inside the phase kernels
    const int *class_rep;            // [n_classes] a component of the class (its layer rows define the table)
    int n_classes;
    double2 *layer_tab;              // [n_classes][sumX] (+ PX_TABLE_PAD entries)
    double *phase_x;                 // [NP][4] derived: mean, abs_scale, ntop
    double *coef;                    // [NP][ncap] derived, index = power of Q
    const int *prob_model;           // [NP] PX_PROB_* (nullptr: every phase carries its W / P in the pool)
    const double *prob_params;       // [NP][PX_PROB_STRIDE] probability parameters of the model
    int *phase_valid;                // [NP]   derived: valid_probs (host flag, or the device model's verdict)
    // derived by px_planes_phase (k_phase_planes_image): planes of a phase that sit at the same z in several components, listed once
    int acap;                        // row stride of the three tables = most atoms of any phase
    // phase images (k_phase_prologue / k_phase_planes_image): everything k_phase_small stages per block, laid out once per generation so that
    // the staging is ONE round of independent loads instead of a chain of six dependent ones (+ three divisions):
    //   [PX_IMG_HDR header words][plane_arg AE][atom_po 2 AE][plane_end AE / 2][useg AE / 2][P R^2][WG G R] ... [coef ncap]
    double *phase_img;               // [NP][img_stride]
    int img_stride, img_ae, img_mcap;   // words per image; acap rounded up to even; words reserved for P and WG
    int img_ncopy;                   // coefficients a block copies (set per launch: the launch's shared-memory capacity)
    int img_sct_off;                 // byte offset of the sine / cosine table inside the block's dynamic shared memory (per launch)
    int *ph_nu;                      // [NP]        number of distinct planes
    double *ph_uz;                   // [NP][acap]  z of distinct plane u
    int *ph_uend;                    // [NP][acap]  one past the last segment of distinct plane u
    int *ph_useg;                    // [NP][acap]  segments ordered by distinct plane: first atom | end atom << 9 | component << 18
    const int *job_pb, *job_spec;    // [NJ]
    const long long *job_out;        // [NJ]
    const double *fractions, *scales, *bgshifts;   // [K][M], [K][S], [K][S]
    double *intensities;             // [K][S][M][Z][X_s]
    double *scaled;                  // same shape (optional)
    double *totals;                  // [K][S][Z][X_s]
    double *residuals;               // [K][1+S]
};

// ---------------------------------------------------------------------------------
// static tables
// ---------------------------------------------------------------------------------
__global__ void k_specimen_derived(int S, const double *gonio, double *derived) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const double *g = gonio + s * 12;
    double mcr = g[3];
    double c = cos(mcr * (PX_PI / 180.0));                 // np.cos(np.radians(mcr)) ** 2
    double h1 = g[1] * 0.5, h2 = g[2] * 0.5;
    double Ss = sqrt(__dadd_rn(__dmul_rn(h1, h1), __dmul_rn(h2, h2)));   // goniometer.py:15-18
    derived[s * 8 + 0] = __dmul_rn(c, c);
    derived[s * 8 + 1] = Ss;
    derived[s * 8 + 2] = __dmul_rn(Ss, Ss);
    // sample-length factor L/(R tan(div)) (specimen.py:41)
    double div = g[5];
    derived[s * 8 + 4] = g[9] / __dmul_rn(g[10], tan(div * (PX_PI / 180.0)));
}

__global__ void k_specimen_tables(StaticDev st, double *sin_t, double *inv_sin, double *stl, double *c2t2, double *corr,
                                  const double *stl_override) {
    const int s = blockIdx.y;
    const int X = st.spec_npts[s];
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= X) return;
    const long long gx = st.spec_off[s] + x;
    const double *g = st.gonio + s * 12;
    const double th = st.theta[gx];
    const double sn = sin(th);
    sin_t[gx] = sn;
    inv_sin[gx] = __ddiv_rn(1.0, sn);
    stl[gx] = stl_override ? stl_override[gx] : __ddiv_rn(__dmul_rn(2.0, sn), g[0]);   // 2 * sin(theta) / wavelength
    const double c2 = cos(__dmul_rn(2.0, th));
    c2t2[gx] = __dmul_rn(c2, c2);
    // machine correction, specimen.py:21-43
    double c = 1.0;
    const int mode = (int)g[4];
    if (mode == 1) c = __dmul_rn(c, sn);                    // AUTOMATIC: * sin(theta)
    if (g[6] > 0.0) {
        const double mu = __dmul_rn(__dmul_rn(g[7], g[8]), 1e-3);
        if (mu > 0.0) c = __dmul_rn(c, fmin(1.0 - exp(__ddiv_rn(__dmul_rn(-2.0, mu), sn)), 1.0));
    }
    if (mode == 0 && g[5] > 0.0) c = __dmul_rn(c, fmin(__dmul_rn(sn, st.derived[s * 8 + 4]), 1.0));
    corr[gx] = c;
}

__device__ __forceinline__ double asf_value(const double *__restrict__ at, double s) {
    // (((a1 e1 + a2 e2) + a3 e3) + a4 e4) + a5 e5, + c, * exp(-B s)   (atoms.py:33-34)
    double acc = __dmul_rn(at[0], exp(__dmul_rn(-at[5], s)));
#pragma unroll
    for (int i = 1; i < 5; ++i) acc = __dadd_rn(acc, __dmul_rn(at[i], exp(__dmul_rn(-at[5 + i], s))));
    acc = __dadd_rn(acc, at[10]);
    return __dmul_rn(acc, exp(__dmul_rn(-at[11], s)));
}

__global__ void k_asf_table(StaticDev st, double *asf) {
    const long long gx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (gx >= st.sumX) return;
    const double v = __dmul_rn(st.stl[gx], 0.05);
    asf[(long long)t * st.sumX + gx] = asf_value(st.atypes + t * 12, __dmul_rn(v, v));
}

__global__ void k_leaf_asf(const double *s, long long n, const double *at, double *out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = asf_value(at, s[i]);
}

// sum |observed| over the selected points of all Z patterns of one specimen (statistics.py:27 denominator)
__global__ void k_observed_norm(StaticDev st, double *derived) {
    const int s = blockIdx.x;
    const int X = st.spec_npts[s];
    const double *obs = st.observed + (long long)st.Z * st.spec_off[s];
    const unsigned char *sel = st.selected + st.spec_off[s];
    double acc = 0.0, cnt = 0.0;
    for (int i = threadIdx.x; i < st.Z * X; i += blockDim.x)
        if (sel[i % X]) { acc += fabs(obs[i]); cnt += 1.0; }
    __shared__ double red[2][32];
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_down_sync(0xffffffffu, acc, o);
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = acc; red[1][threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = threadIdx.x < (blockDim.x >> 5) ? red[0][threadIdx.x] : 0.0;
        cnt = threadIdx.x < (blockDim.x >> 5) ? red[1][threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) {
            acc += __shfl_down_sync(0xffffffffu, acc, o);
            cnt += __shfl_down_sync(0xffffffffu, cnt, o);
        }
        if (threadIdx.x == 0) { derived[s * 8 + 3] = acc; derived[s * 8 + 5] = cnt; }   // sum|obs|, #selected points
    }
}

// ---------------------------------------------------------------------------------
// probability models on the device: W, P and valid_probs of a phase from
// its refinable probability parameters, as pyxrd/probabilities/models do in update() + solve() +
// validate() (base_models.py:142-241).  Every family the reference ships (models/__init__.py:9-12):
// R0 G1..G6 (R0models.py:68-83), R1G2 (R1models.py:125-139), R1G3 (:367-410), R1G4 (:747-797),
// R2G2 (R2models.py:195-233), R2G3 (:485-535), R3G2 (R3models.py:158-198).
// The level matrices lW[r] / lP[r] (rank G^(r+1)) and the multi-index access mW[i,j,..] /
// mP[i,j,..] of base_models.py:52-70,243-285 are kept as they are in the reference, so that
// every model reads like its update(); all arithmetic is single IEEE operations in the
// reference's order (bit-identical W, P; tests/golden/models.npz).  Parameters are clamped to the
// bounds of their properties first (cast_property.py:44-47).  State convention: the converged
// level matrices (a model object that has seen update() twice), i.e. update() is evaluated twice.
// ---------------------------------------------------------------------------------
#define PX_PROB_FIXED 0
#define PX_PROB_R0 1
#define PX_PROB_R1G2 2
#define PX_PROB_R2G2 3
#define PX_PROB_R3G2 4
#define PX_PROB_R1G3 5
#define PX_PROB_R1G4 6
#define PX_PROB_R2G3 7
#define PX_PROB_STRIDE 12
#define PX_PROB_MAXSQ 81  // largest level matrix: rank 9 (R2G3)

// G, R compile-time: every multi-index of a model's update() folds to a constant and the level matrices live in
// registers (the first version kept them in 4.5 KB of local memory per thread behind run-time index arithmetic:
// 51 us for c4's 1536 model-driven phases, 15 us for c2's 2048)
template <int G_, int R_>
struct ProbLevels {
    static constexpr int G = G_, R = R_, Rp = R_ > 1 ? R_ : 1;
    __host__ __device__ static constexpr int ipow(int e) { int v = 1; for (int i = 0; i < e; ++i) v *= G_; return v; }
    static constexpr int SQ = ipow(Rp) * ipow(Rp);  // largest level matrix
    double W[Rp + 1][SQ];            // lW[0 .. Rp-1], lW[Rp] = W . P
    double P[Rp][SQ];                // lP[0 .. Rp-1]
    __device__ __forceinline__ static constexpr int wrank(int r) { return ipow(r < Rp ? r + 1 : Rp); }
    __device__ __forceinline__ static constexpr int prank(int r) { return ipow(r + 1); }
    __device__ __forceinline__ static void xy(const int *ind, int Rl, int &x, int &y) {
        x = 0; y = 0;
        for (int i = 1; i <= Rl; ++i) { const int f = ipow(Rl - i); x += ind[i - 1] * f; y += ind[i] * f; }
    }
    __device__ __forceinline__ double &mW(const int *ind, int l) {             // base_models.py:263-281
        if (l == Rp + 1) {
            int x, y;
            xy(ind, l - 1 > 1 ? l - 1 : 1, x, y);
            return W[l - 1][x * wrank(l - 1) + y];
        }
        int x = 0;
        for (int i = 0; i < l; ++i) x += ind[i] * ipow(l - (i + 1));
        return W[l - 1][x * wrank(l - 1) + x];
    }
    __device__ __forceinline__ double &mP(const int *ind, int l) {             // base_models.py:243-261
        int x, y;
        xy(ind, l - 1 > 1 ? l - 1 : 1, x, y);
        return P[l - 2][x * prank(l - 2) + y];
    }
    __device__ __forceinline__ double &W1(int a) { int i[1] = {a}; return mW(i, 1); }
    __device__ __forceinline__ double &W2(int a, int b) { int i[2] = {a, b}; return mW(i, 2); }
    __device__ __forceinline__ double &W3(int a, int b, int c) { int i[3] = {a, b, c}; return mW(i, 3); }
    __device__ __forceinline__ double &P2(int a, int b) { int i[2] = {a, b}; return mP(i, 2); }
    __device__ __forceinline__ double &P3(int a, int b, int c) { int i[3] = {a, b, c}; return mP(i, 3); }
    __device__ __forceinline__ double &P4(int a, int b, int c, int d) { int i[4] = {a, b, c, d}; return mP(i, 4); }
};

__device__ __forceinline__ double pm_clamp(double v, double lo) { return fmin(fmax(v, lo), 1.0); }
__device__ __forceinline__ double pm_c01(double v) { return fmax(fmin(v, 1.0), 0.0); }      // max(min(v, 1.0), 0.0)
__device__ __forceinline__ double pm_sub(double a, double b) { return __dadd_rn(a, -b); }
__device__ __forceinline__ double pm_muldiv(double a, double b, double c) { return __ddiv_rn(__dmul_rn(a, b), c); }
__device__ __forceinline__ double pm_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double pm_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double pm_div(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double pm_ratio(double a, double b) { return b > 0.0 ? __ddiv_rn(a, b) : 0.0; }   // a / b if b > 0 else 0

template <class LV>
__device__ __forceinline__ void prob_solve(LV &lv) {                    // base_models.py:142-162
    constexpr int G = LV::G;
#pragma unroll
    for (int num = 1; num < LV::R; ++num) {
        int base[4];
#pragma unroll
        for (int code = 0; code < LV::ipow(num); ++code) {
#pragma unroll
            for (int i = 0, c = code; i < num; ++i) { base[num - i] = c % G; c /= G; }      // base[1..num]
            double acc = 0.0;
#pragma unroll
            for (int i = 0; i < G; ++i) { base[0] = i; acc = __dadd_rn(acc, lv.mW(base, num + 1)); }
            lv.mW(base + 1, num) = acc;
        }
#pragma unroll
        for (int code = 0; code < LV::ipow(num + 1); ++code) {
#pragma unroll
            for (int i = 0, c = code; i <= num; ++i) { base[num - i] = c % G; c /= G; }      // base[0..num]
            const double w = lv.mW(base, num);
            lv.mP(base, num + 1) = w > 0.0 ? __ddiv_rn(lv.mW(base, num + 1), w) : 0.0;
        }
    }
    constexpr int r = LV::wrank(LV::Rp);
#pragma unroll
    for (int i = 0; i < r; ++i)
#pragma unroll
        for (int j = 0; j < r; ++j) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < r; ++k) acc = fma(lv.W[LV::Rp - 1][i * r + k], lv.P[LV::Rp - 1][k * r + j], acc);
            lv.W[LV::Rp][i * r + j] = acc;
        }
}

template <int model, class LV>
__device__ __forceinline__ void prob_update(LV &lv, const double *par) {
    constexpr int G = LV::G;
    if constexpr (model == PX_PROB_R0) {
        if (G > 1) {
            double run = 0.0;                                   // np.sum(np.diag(W)[0:i]), sequential
#pragma unroll
            for (int i = 0; i < G - 1; ++i) {
                const double F = pm_clamp(par[i], 0.0);
                const double w = i > 0 ? __dmul_rn(F, pm_sub(1.0, run)) : F;
                lv.W1(i) = w;
                run = i > 0 ? __dadd_rn(run, w) : w;
            }
            lv.W1(G - 1) = pm_sub(1.0, run);
        } else {
            lv.W1(0) = 1.0;
        }
#pragma unroll
        for (int i = 0; i < G; ++i)
#pragma unroll
            for (int j = 0; j < G; ++j) lv.P[0][i * G + j] = lv.W[0][j * G + j];
    } else if constexpr (model == PX_PROB_R1G2) {
        const double W1 = pm_clamp(par[0], 0.0), Pxx = pm_clamp(par[1], 0.0);
        lv.W1(0) = W1;
        lv.W1(1) = pm_sub(1.0, W1);
        if (lv.W1(0) <= 0.5) {
            lv.P2(0, 0) = Pxx;
            lv.P2(0, 1) = pm_sub(1.0, lv.P2(0, 0));
            lv.P2(1, 0) = pm_muldiv(lv.W1(0), lv.P2(0, 1), lv.W1(1));
            lv.P2(1, 1) = pm_sub(1.0, lv.P2(1, 0));
        } else {
            lv.P2(1, 1) = Pxx;
            lv.P2(1, 0) = pm_sub(1.0, lv.P2(1, 1));
            lv.P2(0, 1) = pm_muldiv(lv.W1(1), lv.P2(1, 0), lv.W1(0));
            lv.P2(0, 0) = pm_sub(1.0, lv.P2(0, 1));
        }
    } else if constexpr (model == PX_PROB_R2G2) {
        const double W1 = pm_clamp(par[0], 0.5), Pa = pm_clamp(par[1], 0.0), P21 = pm_clamp(par[2], 0.0),
                     Pb = pm_clamp(par[3], 0.0);
        lv.W1(0) = W1;
        lv.W1(1) = pm_sub(1.0, lv.W1(0));
        lv.P2(1, 0) = P21;
        lv.P2(1, 1) = pm_sub(1.0, lv.P2(1, 0));
        lv.W2(1, 0) = __dmul_rn(lv.W1(1), lv.P2(1, 0));
        lv.W2(1, 1) = __dmul_rn(lv.W1(1), lv.P2(1, 1));
        lv.W2(0, 1) = lv.W2(1, 0);
        lv.W2(0, 0) = pm_sub(lv.W1(0), lv.W2(1, 0));
        if (lv.W1(0) <= 2.0 / 3.0) {
            lv.P3(0, 0, 1) = Pa;
            lv.P3(1, 0, 0) = lv.W2(1, 0) == 0.0 ? 0.0 : pm_muldiv(lv.P3(0, 0, 1), lv.W2(0, 0), lv.W2(1, 0));
        } else {
            lv.P3(1, 0, 0) = Pa;
            lv.P3(0, 0, 1) = lv.W2(0, 0) == 0.0 ? 0.0 : pm_muldiv(lv.P3(1, 0, 0), lv.W2(1, 0), lv.W2(0, 0));
        }
        lv.P3(1, 0, 1) = pm_sub(1.0, lv.P3(1, 0, 0));
        lv.P3(0, 0, 0) = pm_sub(1.0, lv.P3(0, 0, 1));
        if (lv.P2(1, 0) <= 0.5) {
            lv.P3(0, 1, 1) = Pb;
            lv.P3(1, 1, 0) = pm_muldiv(lv.P3(0, 1, 1), lv.W2(0, 1), lv.W2(1, 1));
        } else {
            lv.P3(1, 1, 0) = Pb;
            lv.P3(0, 1, 1) = pm_muldiv(lv.P3(1, 1, 0), lv.W2(1, 1), lv.W2(0, 1));
        }
        lv.P3(0, 1, 0) = pm_sub(1.0, lv.P3(0, 1, 1));
        lv.P3(1, 1, 1) = pm_sub(1.0, lv.P3(1, 1, 0));
    } else if constexpr (model == PX_PROB_R3G2) {
        const double W1 = pm_clamp(par[0], 2.0 / 3.0), Px = pm_clamp(par[1], 0.0);
        lv.W1(0) = W1;
        lv.W1(1) = pm_sub(1.0, W1);
        const double w0 = lv.W1(0), w1 = lv.W1(1), w0m2w1 = pm_sub(w0, __dmul_rn(2.0, w1));
        if (w0 <= 0.75) {
            lv.P4(0, 0, 0, 0) = Px;
            lv.P4(0, 0, 0, 1) = pm_c01(pm_sub(1.0, lv.P4(0, 0, 0, 0)));
            lv.P4(1, 0, 0, 0) = pm_c01(pm_muldiv(lv.P4(0, 0, 0, 1), w0m2w1, w1));
            lv.P4(1, 0, 0, 1) = pm_c01(pm_sub(1.0, lv.P4(1, 0, 0, 0)));
        } else {
            lv.P4(1, 0, 0, 1) = Px;
            lv.P4(1, 0, 0, 0) = pm_c01(pm_sub(1.0, lv.P4(1, 0, 0, 1)));
            lv.P4(0, 0, 0, 0) = pm_c01(pm_sub(1.0, pm_muldiv(lv.P4(1, 0, 0, 0), w1, w0m2w1)));
            lv.P4(0, 0, 0, 1) = pm_c01(pm_sub(1.0, lv.P4(0, 0, 0, 0)));
        }
        const int fixed[6][3] = {{0, 0, 1}, {0, 1, 0}, {0, 1, 1}, {1, 0, 1}, {1, 1, 0}, {1, 1, 1}};
#pragma unroll
        for (int t = 0; t < 6; ++t) {
            lv.P4(fixed[t][0], fixed[t][1], fixed[t][2], 0) = 1.0;
            lv.P4(fixed[t][0], fixed[t][1], fixed[t][2], 1) = 0.0;
        }
        lv.W3(0, 0, 0) = pm_c01(pm_sub(__dmul_rn(3.0, w0), 2.0));
        lv.W3(1, 0, 1) = 0.0; lv.W3(1, 1, 0) = 0.0; lv.W3(1, 1, 1) = 0.0; lv.W3(0, 1, 1) = 0.0;
        const double wm = pm_c01(pm_sub(1.0, w0));
        lv.W3(1, 0, 0) = wm; lv.W3(0, 1, 0) = wm; lv.W3(0, 0, 1) = wm;
    } else if constexpr (model == PX_PROB_R1G3) {
        const double W1 = pm_clamp(par[0], 0.0), Pxx = pm_clamp(par[1], 0.0), G1 = pm_clamp(par[2], 0.0),
                     G2 = pm_clamp(par[3], 0.0), G3 = pm_clamp(par[4], 0.0), G4 = pm_clamp(par[5], 0.0);
        lv.W1(0) = W1;
        lv.W1(1) = pm_mul(pm_sub(1.0, lv.W1(0)), G1);
        lv.W1(2) = pm_sub(pm_sub(1.0, lv.W1(0)), lv.W1(1));
        const double W0inv = lv.W1(0) > 0.0 ? pm_div(1.0, lv.W1(0)) : 0.0;
        double Wxx;
        if (lv.W1(0) <= 0.5) {
            lv.P2(0, 0) = Pxx;
            Wxx = pm_add(pm_add(pm_mul(lv.W1(0), pm_sub(lv.P2(0, 0), 1.0)), lv.W1(1)), lv.W1(2));
        } else {
            Wxx = pm_mul(pm_sub(1.0, lv.W1(0)), Pxx);
        }
        lv.W2(1, 1) = pm_mul(pm_mul(Wxx, G2), G3);
        lv.W2(1, 2) = pm_mul(pm_mul(Wxx, G2), pm_sub(1.0, G3));
        lv.P2(1, 1) = pm_ratio(lv.W2(1, 1), lv.W1(1));
        lv.W2(2, 1) = pm_mul(pm_mul(Wxx, pm_sub(1.0, G2)), G4);
        lv.W2(2, 2) = pm_mul(pm_mul(Wxx, pm_sub(1.0, G2)), pm_sub(1.0, G4));
        lv.P2(1, 2) = pm_ratio(lv.W2(1, 2), lv.W1(1));
        lv.P2(1, 0) = pm_sub(pm_sub(1.0, lv.P2(1, 1)), lv.P2(1, 2));
        lv.P2(2, 1) = pm_ratio(lv.W2(2, 1), lv.W1(2));
        lv.P2(2, 2) = pm_ratio(lv.W2(2, 2), lv.W1(2));
        lv.P2(2, 0) = pm_sub(pm_sub(1.0, lv.P2(2, 1)), lv.P2(2, 2));
        lv.P2(0, 1) = pm_mul(pm_sub(pm_sub(lv.W1(1), lv.W2(1, 1)), lv.W2(2, 1)), W0inv);
        lv.P2(0, 2) = pm_mul(pm_sub(pm_sub(lv.W1(2), lv.W2(1, 2)), lv.W2(2, 2)), W0inv);
        if (lv.W1(0) > 0.5) lv.P2(0, 0) = pm_sub(pm_sub(1.0, lv.P2(0, 1)), lv.P2(0, 2));
        // R1models.py:405-407 rewrite the two-index weights, which solve() replaces by W . P anyway
    } else if constexpr (model == PX_PROB_R1G4) {
        double q[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) q[i] = pm_clamp(par[i], 0.0);      // W1 P11_or_P22 R1 R2 G1 G2 G11 G12 G21 G22 G31 G32
        lv.W1(0) = q[0];
        lv.W1(1) = pm_mul(pm_sub(1.0, lv.W1(0)), q[2]);
        lv.W1(2) = pm_mul(pm_sub(pm_sub(1.0, lv.W1(0)), lv.W1(1)), q[3]);
        lv.W1(3) = pm_sub(pm_sub(pm_sub(1.0, lv.W1(0)), lv.W1(1)), lv.W1(2));
        const double W0inv = lv.W1(0) > 0.0 ? pm_div(1.0, lv.W1(0)) : 0.0;
        double Wxx;
        if (lv.W1(0) < 0.5) {
            lv.P2(0, 0) = q[1];
            Wxx = pm_add(pm_add(pm_add(pm_mul(lv.W1(0), pm_sub(lv.P2(0, 0), 1.0)), lv.W1(1)), lv.W1(2)), lv.W1(3));
        } else {
            Wxx = pm_mul(q[1], pm_add(pm_add(lv.W1(1), lv.W1(2)), lv.W1(3)));
        }
        const double W1x = pm_mul(Wxx, q[4]), Wyx = pm_sub(Wxx, W1x), W2x = pm_mul(Wyx, q[5]), W3x = pm_sub(Wyx, W2x);
        const double Wix[3] = {W1x, W2x, W3x};
#pragma unroll
        for (int i = 1; i < 4; ++i) {
            lv.W2(i, 1) = pm_mul(q[4 + 2 * i], Wix[i - 1]);
            lv.W2(i, 2) = pm_mul(q[5 + 2 * i], pm_sub(Wix[i - 1], lv.W2(i, 1)));
            lv.W2(i, 3) = pm_sub(pm_sub(Wix[i - 1], lv.W2(i, 1)), lv.W2(i, 2));
        }
#pragma unroll
        for (int i = 1; i < 4; ++i) {
            lv.P2(i, 0) = 1.0;
#pragma unroll
            for (int j = 1; j < 4; ++j) {
                lv.P2(i, j) = pm_ratio(lv.W2(i, j), lv.W1(i));
                lv.P2(i, 0) = pm_sub(lv.P2(i, 0), lv.P2(i, j));
            }
            lv.W2(i, 0) = pm_mul(lv.W1(i), lv.P2(i, 0));
        }
#pragma unroll
        for (int j = 1; j < 4; ++j)
            lv.P2(0, j) = pm_mul(pm_sub(pm_sub(pm_sub(lv.W1(j), lv.W2(1, j)), lv.W2(2, j)), lv.W2(3, j)), W0inv);
        if (lv.W1(0) >= 0.5) lv.P2(0, 0) = pm_sub(pm_sub(pm_sub(1.0, lv.P2(0, 1)), lv.P2(0, 2)), lv.P2(0, 3));
    } else if constexpr (model == PX_PROB_R2G3) {
        const double W1 = pm_clamp(par[0], 0.5), Pxx = pm_clamp(par[1], 0.0), G1 = pm_clamp(par[2], 0.0),
                     G2 = pm_clamp(par[3], 0.0), G3 = pm_clamp(par[4], 0.0), G4 = pm_clamp(par[5], 0.0);
        lv.W1(0) = W1;
        lv.W1(1) = pm_mul(pm_sub(1.0, lv.W1(0)), G1);
        lv.W1(2) = pm_sub(pm_sub(1.0, lv.W1(0)), lv.W1(1));
        lv.W2(1, 1) = 0.0; lv.W2(1, 2) = 0.0; lv.W2(2, 1) = 0.0; lv.W2(2, 2) = 0.0;
        lv.W2(1, 0) = lv.W1(1); lv.W2(0, 1) = lv.W1(1); lv.W3(0, 1, 0) = lv.W1(1);
        lv.W2(2, 0) = lv.W1(2); lv.W2(0, 2) = lv.W1(2); lv.W3(0, 2, 0) = lv.W1(2);
        lv.W2(0, 0) = pm_sub(pm_sub(lv.W1(0), lv.W2(0, 1)), lv.W2(0, 2));
        const double Wx = pm_add(lv.W1(1), lv.W1(2));
        double Px0x;
        if (lv.W1(0) < 2.0 / 3.0) {
            lv.P3(0, 0, 0) = Pxx;
            Px0x = Wx != 0.0 ? pm_sub(1.0, pm_mul(pm_div(pm_sub(lv.W1(0), Wx), Wx), pm_sub(1.0, lv.P3(0, 0, 0)))) : 0.0;
        } else {
            Px0x = Pxx;
            const double den = pm_sub(lv.W1(0), Wx);
            lv.P3(0, 0, 0) = den != 0.0 ? pm_sub(1.0, pm_mul(pm_div(Wx, den), pm_sub(1.0, Px0x))) : 0.0;
        }
        const double Wx0x = pm_mul(Wx, Px0x), W10x = pm_mul(G2, Wx0x), W20x = pm_sub(Wx0x, W10x);
        lv.W3(1, 0, 1) = pm_mul(G3, W10x);
        lv.W3(1, 0, 2) = pm_mul(pm_sub(1.0, G3), W10x);
        lv.W3(1, 0, 0) = pm_sub(pm_sub(lv.W2(1, 0), lv.W3(1, 0, 1)), lv.W3(1, 0, 2));
        lv.W3(2, 0, 1) = pm_mul(G4, W20x);
        lv.W3(2, 0, 2) = pm_mul(pm_sub(1.0, G4), W20x);
        lv.W3(2, 0, 0) = pm_sub(pm_sub(lv.W2(2, 0), lv.W3(2, 0, 1)), lv.W3(2, 0, 2));
        lv.W3(0, 0, 0) = pm_mul(lv.W2(0, 0), lv.P3(0, 0, 0));
        lv.W3(0, 0, 1) = pm_sub(pm_sub(lv.W2(0, 1), lv.W3(1, 0, 1)), lv.W3(2, 0, 1));
        lv.W3(0, 0, 2) = pm_sub(pm_sub(lv.W2(0, 2), lv.W3(1, 0, 2)), lv.W3(2, 0, 2));
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j)
#pragma unroll
                for (int k = 0; k < 3; ++k) lv.P3(i, j, k) = pm_ratio(lv.W3(i, j, k), lv.W2(i, j));
    }
    prob_solve(lv);
}

// -> rank; W (diagonal matrix) and P written row-major, *valid = every entry of every level in [0, 1]
template <int model, int G, int R>
__device__ __noinline__ int prob_evaluate_t(const double *par, double *Wout, double *Pout, int *valid) {
    using LV = ProbLevels<G, R>;
    LV lv;
#pragma unroll
    for (int r = 0; r <= LV::Rp; ++r)
#pragma unroll
        for (int i = 0; i < LV::SQ; ++i) lv.W[r][i] = 0.0;
#pragma unroll
    for (int r = 0; r < LV::Rp; ++r)
#pragma unroll
        for (int i = 0; i < LV::SQ; ++i) lv.P[r][i] = 0.0;
    prob_update<model>(lv, par);
    prob_update<model>(lv, par);
    bool ok = true;
#pragma unroll
    for (int r = 0; r <= LV::Rp; ++r)
#pragma unroll
        for (int i = 0; i < LV::wrank(r) * LV::wrank(r); ++i) ok = ok && !(lv.W[r][i] < 0.0 || lv.W[r][i] > 1.0);
#pragma unroll
    for (int r = 0; r < LV::Rp; ++r)
#pragma unroll
        for (int i = 0; i < LV::prank(r) * LV::prank(r); ++i) ok = ok && !(lv.P[r][i] < 0.0 || lv.P[r][i] > 1.0);
    constexpr int rank = LV::wrank(LV::Rp);
#pragma unroll
    for (int i = 0; i < rank * rank; ++i) { Wout[i] = lv.W[LV::Rp - 1][i]; Pout[i] = lv.P[LV::Rp - 1][i]; }
    *valid = ok ? 1 : 0;
    return rank;
}

__device__ int prob_evaluate(int model, int G, const double *par, double *Wout, double *Pout, int *valid) {
    switch (model) {
        case PX_PROB_R0:
            switch (G) {
                case 1: return prob_evaluate_t<PX_PROB_R0, 1, 0>(par, Wout, Pout, valid);
                case 2: return prob_evaluate_t<PX_PROB_R0, 2, 0>(par, Wout, Pout, valid);
                case 3: return prob_evaluate_t<PX_PROB_R0, 3, 0>(par, Wout, Pout, valid);
                case 4: return prob_evaluate_t<PX_PROB_R0, 4, 0>(par, Wout, Pout, valid);
                case 5: return prob_evaluate_t<PX_PROB_R0, 5, 0>(par, Wout, Pout, valid);
                default: return prob_evaluate_t<PX_PROB_R0, 6, 0>(par, Wout, Pout, valid);
            }
        case PX_PROB_R1G2: return prob_evaluate_t<PX_PROB_R1G2, 2, 1>(par, Wout, Pout, valid);
        case PX_PROB_R2G2: return prob_evaluate_t<PX_PROB_R2G2, 2, 2>(par, Wout, Pout, valid);
        case PX_PROB_R3G2: return prob_evaluate_t<PX_PROB_R3G2, 2, 3>(par, Wout, Pout, valid);
        case PX_PROB_R1G3: return prob_evaluate_t<PX_PROB_R1G3, 3, 1>(par, Wout, Pout, valid);
        case PX_PROB_R1G4: return prob_evaluate_t<PX_PROB_R1G4, 4, 1>(par, Wout, Pout, valid);
        default: return prob_evaluate_t<PX_PROB_R2G3, 3, 2>(par, Wout, Pout, valid);
    }
}

// ---------------------------------------------------------------------------------
// population from a template + binding table, expanded on the device:
// per generation only the solution vectors x[K][d] cross PCIe.  Block k writes candidate k's slice of
// every batch array (template values, table indices moved into the candidate's own tables), applies the
// bindings slot = scale * x[k][param] + offset in table order (a later row wins, as in the host version
// pyxrd_expand_population), then the automatic CSDS maximum int(factor * average)
// (phases/models/CSDS.py:127,140).  Job tables and the per-class job lists are generated alongside.
// ---------------------------------------------------------------------------------
struct ExpandDev {
    int K, D, NB, NP, NC, NA, NPC, M, S, JPC, NCLS;
    long long NM, cand_stride;
    // template (one candidate)
    const int *t_phase_i, *t_phase_comp, *t_comp_i, *t_atom_t, *t_job_pb, *t_job_spec, *t_jobids, *t_pmodel, *t_comp_class;
    const long long *t_job_out;
    const double *t_phase_d, *t_matrices, *t_comp_d, *t_atom_d, *t_frac, *t_scales, *t_bg, *t_pparam;
    const int *cls_first, *cls_n;            // [NCLS] position / length of each class inside t_jobids
    const double *x;                         // [K][D]
    const pyxrd_binding *bind;               // [NB]
    const unsigned char *auto_max;           // [NP] or nullptr
    double factor;
    // destination (K candidates)
    int *phase_i, *phase_comp, *comp_i, *atom_t, *job_pb, *job_spec, *jobids, *pmodel, *comp_class;
    long long *job_out;
    double *phase_d, *matrices, *comp_d, *atom_d, *frac, *scales, *bg, *pparam;
};

__global__ void __launch_bounds__(128) k_expand_population(ExpandDev e) {
    const int k = blockIdx.x, t = threadIdx.x, nt = blockDim.x;
    for (int i = t; i < e.NP * 8; i += nt) {
        int v = e.t_phase_i[i];
        if ((i & 7) == 5) v += k * e.NPC;
        if ((i & 7) == 6) v += (int)(k * e.NM);
        e.phase_i[(size_t)k * e.NP * 8 + i] = v;
    }
    for (int i = t; i < e.NP * 8; i += nt) e.phase_d[(size_t)k * e.NP * 8 + i] = e.t_phase_d[i];
    for (int i = t; i < e.NPC; i += nt) e.phase_comp[(size_t)k * e.NPC + i] = e.t_phase_comp[i] + k * e.NC;
    for (long long i = t; i < e.NM; i += nt) e.matrices[(size_t)k * e.NM + i] = e.t_matrices[i];
    for (int i = t; i < e.NC * 8; i += nt) e.comp_d[(size_t)k * e.NC * 8 + i] = e.t_comp_d[i];
    for (int i = t; i < e.NC * 4; i += nt) e.comp_i[(size_t)k * e.NC * 4 + i] = e.t_comp_i[i] + ((i & 3) == 0 ? k * e.NA : 0);
    for (int i = t; i < e.NC; i += nt) e.comp_class[(size_t)k * e.NC + i] = e.t_comp_class[i];
    for (int i = t; i < e.NA * 2; i += nt) e.atom_d[(size_t)k * e.NA * 2 + i] = e.t_atom_d[i];
    for (int i = t; i < e.NA; i += nt) e.atom_t[(size_t)k * e.NA + i] = e.t_atom_t[i];
    for (int i = t; i < e.M; i += nt) e.frac[(size_t)k * e.M + i] = e.t_frac[i];
    for (int i = t; i < e.S; i += nt) { e.scales[(size_t)k * e.S + i] = e.t_scales[i]; e.bg[(size_t)k * e.S + i] = e.t_bg[i]; }
    if (e.pmodel) {
        for (int i = t; i < e.NP; i += nt) e.pmodel[(size_t)k * e.NP + i] = e.t_pmodel[i];
        for (int i = t; i < e.NP * PX_PROB_STRIDE; i += nt) e.pparam[(size_t)k * e.NP * PX_PROB_STRIDE + i] = e.t_pparam[i];
    }
    for (int i = t; i < e.JPC; i += nt) {
        const int pb = e.t_job_pb[i];
        e.job_pb[(size_t)k * e.JPC + i] = pb >= 0 ? pb + k * e.NP : pb;
        e.job_spec[(size_t)k * e.JPC + i] = e.t_job_spec[i];
        e.job_out[(size_t)k * e.JPC + i] = e.t_job_out[i] + k * e.cand_stride;
    }
    for (int c = 0; c < e.NCLS; ++c) {
        const int first = e.cls_first[c], n = e.cls_n[c];
        for (int i = t; i < n; i += nt) e.jobids[(size_t)first * e.K + (size_t)k * n + i] = k * e.JPC + e.t_jobids[first + i];
    }
    __syncthreads();
    if (t == 0) {
        const double *xk = e.x + (size_t)k * e.D;
        for (int b = 0; b < e.NB; ++b) {
            const pyxrd_binding bd = e.bind[b];
            const double v = __dadd_rn(__dmul_rn(bd.scale, xk[bd.param]), bd.offset);
            double *slot;
            switch (bd.target) {
                case 0: slot = e.phase_d + (size_t)k * e.NP * 8; break;
                case 1: slot = e.comp_d + (size_t)k * e.NC * 8; break;
                case 2: slot = e.atom_d + (size_t)k * e.NA * 2; break;
                case 3: slot = e.matrices + (size_t)k * e.NM; break;
                case 4: slot = e.pparam + (size_t)k * e.NP * PX_PROB_STRIDE; break;
                case 5: slot = e.frac + (size_t)k * e.M; break;
                case 6: slot = e.scales + (size_t)k * e.S; break;
                default: slot = e.bg + (size_t)k * e.S; break;
            }
            slot += bd.index;
            *slot = bd.mode == 1 ? __dmul_rn(*slot, v) : bd.mode == 2 ? __dadd_rn(*slot, v) : v;
        }
    }
    __syncthreads();
    if (e.auto_max)
        for (int p = t; p < e.NP; p += nt)
            if (e.auto_max[p])
                e.phase_i[((size_t)k * e.NP + p) * 8 + 3] = (int)__dmul_rn(e.factor, e.phase_d[((size_t)k * e.NP + p) * 8 + 1]);
}

// ---------------------------------------------------------------------------------
// per-candidate prologue
// ---------------------------------------------------------------------------------
__device__ __forceinline__ double lognormal_dev(double T, double a, double b) {
    // exp(-(log(T) - a) ** 2 / (2.0 * b ** 2)) / (sqrt2pi * abs(b) * T)     math_tools.py:36-37
    const double d = log(T) - a;
    return exp(-(d * d) / (2.0 * (b * b))) / (PX_SQRT2PI * fabs(b) * T);
}

// q_T (normalised) into arr[0..maximum], mean, and coefficient-by-power table cpow[0..maximum+1]
__device__ void csds_series(const double *pd, int nmin, int nmax, double *arr, double *mean_out,
                            double *cpow, int *ntop_out) {
    const double lavg = log(pd[1]);
    const double a = pd[2] * lavg + pd[3];
    const double b = sqrt(pd[4] * lavg + pd[5]);
    double total = 0.0, mean = 0.0;
    for (int T = nmin; T <= nmax; ++T) {                   // CSDS.py:29-36,40-44: key int(T), T = max(min+i, 1e-50)
        const double q = lognormal_dev(fmax((double)T, 1e-50), a, b);
        total += q;
        mean += (double)T * q;
    }
    *mean_out = mean / total;
    if (arr) {
        for (int T = 0; T <= nmax; ++T) arr[T] = 0.0;
        for (int T = (nmin > 0 ? nmin : 0); T <= nmax; ++T) arr[T] = lognormal_dev(fmax((double)T, 1e-50), a, b) / total;
    }
    // coef_n = 2 sum_{m>n} (m-n) q_m (phases.py:147-151) by the two running sums
    //   tail(n) = sum_{m>n} q_m,  c(n-1) = c(n) + tail(n-1)     (all terms positive: no cancellation)
    for (int n = 0; n <= nmax + 1; ++n) cpow[n] = 0.0;
    double tail = 0.0, c = 0.0;
    int ntop = 0;
    for (int n = nmax; n >= 0; --n) {
        // here: tail = sum_{m>n} q_m, c = sum_{m>n} (m-n) q_m
        if (n >= nmin) {
            const int power = (n >= 1) ? n : nmax + 1;     // Qn[n-1] with python wrap-around for n = 0
            cpow[power] += 2.0 * c;
            if (c != 0.0 && power > ntop) ntop = power;
        }
        const double qn = (n >= nmin && n >= 1) ? lognormal_dev((double)n, a, b) / total : 0.0;
        tail += qn;
        c += tail;
    }
    *ntop_out = ntop;
}

// thread per phase: W | P of the pool entry and valid_probs from the probability parameters
__global__ void k_probabilities(BatchDev bt) {
    const int pb = blockIdx.x * blockDim.x + threadIdx.x;
    if (pb >= bt.NP) return;
    const int *pi = bt.phase_i + pb * 8;
    int valid = (pi[4] & 1) ? 1 : 0;
    const int model = bt.prob_model ? bt.prob_model[pb] : PX_PROB_FIXED;
    if (valid && model != PX_PROB_FIXED && !(pi[4] & 8)) {
        double *W = const_cast<double *>(bt.matrices) + pi[6];
        prob_evaluate(model, pi[0], bt.prob_params + pb * PX_PROB_STRIDE, W, W + pi[1] * pi[1], &valid);
    }
    bt.phase_valid[pb] = valid;
}

__global__ void k_leaf_probabilities(int model, int G, const double *par, double *W, double *P, int *out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int valid;
    out[1] = prob_evaluate(model, G, par, W, P, &valid);
    out[0] = valid;
}

// warp per phase.  The transcendental part of csds_series (one log-normal density per layer count) is spread over
// the lanes; the sums are then taken by lane 0 in the order csds_series takes them, so both give the same bits.
#define PX_PROLOGUE_WARPS 4
#define PX_PROLOGUE_QCAP 512      // layer counts staged per warp; longer distributions take the serial path
// coefficient half of the phase image (see k_phase_planes_image for the other half and the layout): mean / absolute
// scale / ntop / flags, P, WG, coefficients -- by the warp that has just computed them
__device__ __forceinline__ void px_image_coef(const BatchDev &bt, const int pb, const int lane) {
    const int *pi = bt.phase_i + pb * 8;
    if (!(pi[4] & 1) || (pi[4] & 8)) return;                 // no components / raw pattern: no phase kernel reads an image
    const int G = pi[0], R = pi[1];
    double *img = bt.phase_img + (long long)pb * bt.img_stride;
    if (lane < 3) img[lane] = bt.phase_x[pb * 4 + lane];
    if (lane == 3) img[3] = (double)pi[4];
    double *mat = img + PX_IMG_HDR + 4 * bt.img_ae;
    if (R * R + G * R <= bt.img_mcap) {
        const double *W = bt.matrices + pi[6];
        const double *P = W + R * R;
        const int REPS = R / G;
        for (int i = lane; i < R * R; i += 32) mat[i] = P[i];
        for (int i = lane; i < G * R; i += 32) {             // WG[c][K] = sum_{J in c} W[J][K]
            const int c = i / R, K = i % R;
            double acc = 0.0;
            for (int J = c * REPS; J < (c + 1) * REPS; ++J) acc += W[J * R + K];
            mat[R * R + i] = acc;
        }
    }
    double *cf = mat + bt.img_mcap;
    const double *coef = bt.coef + (long long)pb * bt.ncap;
    for (int i = lane; i < bt.ncap; i += 32) cf[i] = coef[i];
}

__global__ void __launch_bounds__(32 * PX_PROLOGUE_WARPS) k_phase_prologue(BatchDev bt) {
    __shared__ double sq[PX_PROLOGUE_WARPS][PX_PROLOGUE_QCAP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pb = blockIdx.x * PX_PROLOGUE_WARPS + warp;
    if (pb >= bt.NP) return;
    const int *pi = bt.phase_i + pb * 8;
    const double *pd = bt.phase_d + pb * 8;
    double *px = bt.phase_x + pb * 4;
    if (!bt.phase_valid[pb] || (pi[4] & 8)) {               // invalid probabilities / raw pattern
        if (lane == 0) { px[0] = 0.0; px[1] = 0.0; px[2] = 0.0; }
        return;
    }
    const int G = pi[0], r = pi[1], nmin = pi[2], nmax = pi[3];
    double *cpow = bt.coef + (long long)pb * bt.ncap;
    double mean;
    int ntop;
    if (nmax - nmin + 1 > PX_PROLOGUE_QCAP) {
        if (lane == 0) csds_series(pd, nmin, nmax, nullptr, &mean, cpow, &ntop);
    } else {
        double *q = sq[warp];
        const double lavg = log(pd[1]);
        const double a = pd[2] * lavg + pd[3];
        const double b = sqrt(pd[4] * lavg + pd[5]);
        for (int T = nmin + lane; T <= nmax; T += 32) q[T - nmin] = lognormal_dev(fmax((double)T, 1e-50), a, b);
        for (int n = lane; n <= nmax + 1; n += 32) cpow[n] = 0.0;
        __syncwarp();
        if (lane == 0) {
            double total = 0.0, msum = 0.0;
            for (int T = nmin; T <= nmax; ++T) {           // CSDS.py:29-36,40-44
                total += q[T - nmin];
                msum += (double)T * q[T - nmin];
            }
            mean = msum / total;
            // coef_n = 2 sum_{m>n} (m-n) q_m (phases.py:147-151) by the two running sums of csds_series
            double tail = 0.0, c = 0.0;
            ntop = 0;
            for (int n = nmax; n >= 0; --n) {
                if (n >= nmin) {
                    const int power = (n >= 1) ? n : nmax + 1;     // Qn[n-1] with python wrap-around for n = 0
                    cpow[power] += 2.0 * c;
                    if (c != 0.0 && power > ntop) ntop = power;
                }
                const double qn = (n >= nmin && n >= 1) ? q[n - nmin] / total : 0.0;
                tail += qn;
                c += tail;
            }
        }
    }
    if (lane == 0) {
        // absolute scale, phases.py:48-66 (first G diagonal entries of W)
        const double *W = bt.matrices + pi[6];
        double vol = 0.0, d001 = 0.0, dens = 0.0;
        for (int i = 0; i < G; ++i) {
            const double *cd = bt.comp_d + bt.phase_comp[pi[5] + i] * 8;
            const double w = W[i * r + i];
            vol += cd[4] * w;
            d001 += cd[0] * w;
            dens += cd[5] * w / cd[4];
        }
        const double mass = mean * (vol * vol) * dens;
        px[0] = mean;
        px[1] = (mass != 0.0) ? d001 / mass : 0.0;
        px[2] = (double)ntop;
    }
    __syncwarp();
    px_image_coef(bt, pb, lane);                             // (coefficients and phase_x were written by this warp)
}

__device__ __forceinline__ void px_component_z(const BatchDev &bt, const int ci) {
    const double *cd = bt.comp_d + ci * 8;
    const int *cc = bt.comp_i + ci * 4;
    const double lattice_d = cd[3];
    const double zf = __ddiv_rn(cd[0] - lattice_d, cd[1] - lattice_d);   // components.py:36
    int plane = -1;
    double prev = 0.0;
    for (int a = 0; a < cc[1] + cc[2]; ++a) {
        const double z0 = bt.atom_d[(cc[0] + a) * 2 + 1];
        const double z = (a < cc[1]) ? z0 : __dadd_rn(lattice_d, __dmul_rn(zf, z0 - lattice_d));
        bt.atom_z[cc[0] + a] = z;
        if (a == 0 || a == cc[1] || z != prev) ++plane;       // consecutive atoms of one group at one z share a plane
        if (a == cc[1]) bt.comp_ilplane[ci] = plane;
        bt.atom_plane[cc[0] + a] = plane;
        prev = z;
    }
    bt.comp_nplanes[ci] = plane + 1;
    if (cc[2] == 0) bt.comp_ilplane[ci] = plane + 1;
}

// layer table of class blockIdx.y at point gx: (sum over the planes of the layer) (sum_a asf_a pn_a) exp(i (2 pi z) stl),
// the arithmetic the phase kernels apply per plane (plane_term)
__global__ void __launch_bounds__(128) k_layer_table(StaticDev st, BatchDev bt) {
    const long long gx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int cls = blockIdx.y;
    if (gx >= st.sumX) return;
    const int ci = bt.class_rep[cls];
    const int a0 = bt.comp_i[ci * 4], nl = bt.comp_i[ci * 4 + 1];
    const double stl = st.stl[gx];
    double sr = 0.0, si = 0.0;
    int a = 0;
    while (a < nl) {
        const double z = bt.atom_d[(a0 + a) * 2 + 1];
        double f = 0.0;
        for (; a < nl && bt.atom_d[(a0 + a) * 2 + 1] == z; ++a) {
            const int t = bt.atom_type[a0 + a];
            if (t >= 0) f = fma(__ldg(st.asf + (long long)t * st.sumX + gx), bt.atom_d[(a0 + a) * 2], f);
        }
        double sn, cs;
        px_sincos(__dmul_rn(__dmul_rn(PX_TWO_PI, z), stl), sn, cs);
        sr = fma(f, cs, sr);
        si = fma(f, sn, si);
    }
    bt.layer_tab[(long long)cls * st.sumX + gx] = make_double2(sr, si);
}

#define PX_MAX_ATOMS 256     // atoms per phase (all components)

// warp per phase: planes (runs of atoms with equal z inside a component) that occur at the same z in several
// components of the phase -- the shared silicate layer of a mixed-layer phase -- are listed once, so that the phase
// kernels evaluate exp(2 pi i z stl) once per DISTINCT plane.  Segment = (component, atom range) of one plane.
struct PlanesSmem {
    double seg_z[PX_MAX_ATOMS];
    int seg_info[PX_MAX_ATOMS], seg_first[PX_MAX_ATOMS], seg_u[PX_MAX_ATOMS];
};
__device__ __forceinline__ void px_planes_phase(const BatchDev &bt, const int pb, const int lane, PlanesSmem *sm) {
    double *seg_z = sm->seg_z;
    int *seg_info = sm->seg_info, *seg_first = sm->seg_first, *seg_u = sm->seg_u;
    const int *pi = bt.phase_i + pb * 8;
    int *nu_out = bt.ph_nu + pb;
    if (!(pi[4] & 1) || (pi[4] & 8)) { if (lane == 0) *nu_out = 0; return; }
    const int G = pi[0];
    int first_atom = 0, nseg = 0;
    for (int c = 0; c < G; ++c) {
        const int ci = bt.phase_comp[pi[5] + c];
        const int a0 = bt.comp_i[ci * 4], na = bt.comp_i[ci * 4 + 1] + bt.comp_i[ci * 4 + 2];
        const int skip = bt.comp_class[ci] >= 0 ? bt.comp_ilplane[ci] : 0;    // tabulated layer: interlayer planes only
        for (int a = lane; a < na; a += 32) {
            const int pl = bt.atom_plane[a0 + a];
            if (pl >= skip && (a == 0 || bt.atom_plane[a0 + a - 1] != pl)) {
                int e = a + 1;
                while (e < na && bt.atom_plane[a0 + e] == pl) ++e;
                seg_z[nseg + pl - skip] = bt.atom_z[a0 + a];
                seg_info[nseg + pl - skip] = (first_atom + a) | ((first_atom + e) << 9) | (c << 18);
            }
        }
        first_atom += na;
        nseg += bt.comp_nplanes[ci] - skip;
    }
    __syncwarp();
    for (int j = lane; j < nseg; j += 32) {            // first segment at the same z
        int f = j;
        const double z = seg_z[j];
        for (int i = 0; i < j; ++i)
            if (seg_z[i] == z) { f = i; break; }
        seg_first[j] = f;
    }
    __syncwarp();
    for (int j = lane; j < nseg; j += 32) {            // index of the distinct plane a segment belongs to
        const int f = seg_first[j];
        int u = 0;
        for (int i = 0; i < f; ++i) u += (seg_first[i] == i);
        seg_u[j] = u;
    }
    __syncwarp();
    int nu = 0;
    for (int i = 0; i < nseg; ++i) nu += (seg_first[i] == i);
    double *uz = bt.ph_uz + (long long)pb * bt.acap;
    int *uend = bt.ph_uend + (long long)pb * bt.acap, *useg = bt.ph_useg + (long long)pb * bt.acap;
    for (int j = lane; j < nseg; j += 32) {            // stable ordering of the segments by distinct plane
        const int u = seg_u[j];
        int pos = 0, upto = 0;
        for (int i = 0; i < nseg; ++i) {
            const int ui = seg_u[i];
            pos += (ui < u) || (ui == u && i < j);
            upto += (ui <= u);
        }
        useg[pos] = seg_info[j];
        if (seg_first[j] == j) { uz[u] = seg_z[j]; uend[u] = upto; }
    }
    if (lane == 0) *nu_out = nu;
}

__global__ void k_leaf_csds(const double *pd, int nmin, int nmax, double *arr, double *mean, double *cpow) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int ntop;
        csds_series(pd, nmin, nmax, arr, mean, cpow, &ntop);
    }
}

// ---------------------------------------------------------------------------------
// phase kernels: shared-memory staging of one phase parameter block
// ---------------------------------------------------------------------------------

struct PhaseSmem {          // header living at the start of dynamic shared memory
    // ---- header block: PX_IMG_HDR words, copied verbatim from the phase image (k_phase_planes_image) by k_phase_small ----
    double px[4];           // mean, absolute scale, ntop, phase flags (phase_x + phase_i[4])
    double lpf_all[8][4];   // per specimen: lpf_q, lpf_k1, lpf_k2, lpf_pol (lpf_consts)
    double comp_d001[8];
    double comp_delta[8];
    long long comp_tab[8];  // offset of the component's layer class inside the layer table, -1: no table (all planes listed)
    int comp_first[8];      // first plane (local index) of component c
    int comp_count[8];      // number of planes of component c
    int comp_a0[8];         // first atom (local index) of the first listed plane of component c
    int n_atoms, n_u, pad[2];
    // ---- per block ----
    double lpf_q, lpf_k1, lpf_k2, lpf_pol;   // Lorentz-polarisation constants of this (phase, specimen), see lpf_point
    double plane_arg[PX_MAX_ATOMS];  // (2 pi) * z of a plane = run of consecutive atoms with equal z
    int plane_end[PX_MAX_ATOMS];     // one past the last atom (local index) of the plane
    double2 atom_po[PX_MAX_ATOMS];   // .x = pn (0 for a None atom type), .y = bit pattern of the ASF-table offset type * sumX
    int useg[PX_MAX_ATOMS];          // distinct-plane mode (k_phase_small): plane_arg / plane_end describe DISTINCT planes of the
                                     // whole phase, plane_end[u] = one past the last entry of useg that belongs to plane u
};
static_assert(offsetof(PhaseSmem, lpf_q) == PX_IMG_HDR * 8, "header block of PhaseSmem and PX_IMG_HDR disagree");

// Lorentz-polarisation factor for one point (goniometer.py:20-34)
__device__ __forceinline__ double lpf_value(double st, double c2t2, double sigma_star, double Ss, double Ss2, double pol) {
    const double sig = fmax(sigma_star, 1e-18);
    const double Q = Ss / ((PX_SQRT8 * st) * sig);
    const double T = erf(Q) * PX_SQRT2PI / ((2.0 * sig) * Ss) - (2.0 * st) * (1.0 - exp(-(Q * Q))) / Ss2;
    return T * (1.0 + pol * c2t2) / st;
}

// The same factor inside the phase kernels: the divisions by block-uniform quantities are
// multiplications by reciprocals taken once per block, and 1 / sin(theta) is a static table
// (differs from lpf_value by rounding only, ~2 ulp).
//   Q = lpf_q / st, lpf_q = S / (sqrt8 sigma);  T = erf(Q) lpf_k1 - 2 st (1 - exp(-Q^2)) lpf_k2
__device__ __forceinline__ void lpf_consts(double sigma_star, const double *derived_s, double *out4) {
    const double sig = fmax(sigma_star, 1e-18), Ss = derived_s[1], Ss2 = derived_s[2];
    out4[0] = Ss / (PX_SQRT8 * sig);
    out4[1] = PX_SQRT2PI / ((2.0 * sig) * Ss);
    out4[2] = 1.0 / Ss2;
    out4[3] = derived_s[0];
}
__device__ __forceinline__ void lpf_setup(PhaseSmem *ps, double sigma_star, const double *derived_s) {
    lpf_consts(sigma_star, derived_s, &ps->lpf_q);
}
// erf(Q) and exp(-Q^2) come from the interval tables above (absolute error 1.5e-16 / 3.7e-16); the N points of a thread
// run in lock step, neighbouring points fall into the same or the next interval (one cache line per warp and step)
template <int N>
__device__ __forceinline__ void lpf_point_n(const PhaseSmem *ps, const double (&st)[N], const double (&inv_st)[N],
                                            const double (&c2t2)[N], double (&out)[N]) {
    const double2 *row[N];
    double s[N], e[N], g[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double Q = fmin(ps->lpf_q * inv_st[k], (double)PX_LPF_INTERVALS / PX_LPF_WIDTH_INV);
        const int i = min(__double2int_rz(Q * PX_LPF_WIDTH_INV), PX_LPF_INTERVALS);
        s[k] = fma(Q, 2.0 * PX_LPF_WIDTH_INV, -(double)(2 * i + 1));
        row[k] = reinterpret_cast<const double2 *>(PX_LPF_TAB) + i * (PX_LPF_DEGREE + 1);
        const double2 c = __ldg(row[k]);
        e[k] = c.x;
        g[k] = c.y;
    }
#pragma unroll
    for (int j = 1; j <= PX_LPF_DEGREE; ++j)
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const double2 c = __ldg(row[k] + j);
            e[k] = fma(e[k], s[k], c.x);
            g[k] = fma(g[k], s[k], c.y);
        }
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double T = e[k] * ps->lpf_k1 - (2.0 * st[k]) * (1.0 - g[k]) * ps->lpf_k2;
        out[k] = T * (1.0 + ps->lpf_pol * c2t2[k]) * inv_st[k];
    }
}
__device__ __forceinline__ double lpf_point(const PhaseSmem *ps, double st, double inv_st, double c2t2) {
    const double a[1] = {st}, b[1] = {inv_st}, c[1] = {c2t2};
    double out[1];
    lpf_point_n<1>(ps, a, b, c, out);
    return out[0];
}

// t_only: get_T (goniometer.py:20-25) instead of the whole factor
__global__ void k_leaf_lpf(const double *theta, long long n, double sigma_star, double s1, double s2,
                           double mcr, double *out, int t_only) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double c = cos(mcr * (PX_PI / 180.0));
    const double h1 = s1 * 0.5, h2 = s2 * 0.5;
    const double Ss = sqrt(__dadd_rn(__dmul_rn(h1, h1), __dmul_rn(h2, h2)));
    const double th = theta[i];
    const double c2 = cos(2.0 * th);
    if (t_only) {
        const double st = sin(th), sig = fmax(sigma_star, 1e-18);
        const double Q = Ss / ((PX_SQRT8 * st) * sig);
        out[i] = erf(Q) * PX_SQRT2PI / ((2.0 * sig) * Ss) - (2.0 * st) * (1.0 - exp(-(Q * Q))) / __dmul_rn(Ss, Ss);
        return;
    }
    out[i] = lpf_value(sin(th), __dmul_rn(c2, c2), sigma_star, Ss, __dmul_rn(Ss, Ss), __dmul_rn(c, c));
}

// cooperative load of the component / atom lists of phase `pb` into shared memory.  Atoms of a
// component that sit at the same z (several species on one crystallographic plane) share one
// complex exponential: sum_a asf_a pn_a exp(i arg) = (sum_a asf_a pn_a) exp(i arg).
template <bool DISTINCT = false>
__device__ __forceinline__ void load_phase_lists(const BatchDev &bt, int pb, int G, PhaseSmem *ps, long long sumX,
                                                 const double *derived_s) {
    const int *pi = bt.phase_i + pb * 8;
    int first_atom = 0, first_plane = 0;
    for (int c = 0; c < G; ++c) {
        const int ci = bt.phase_comp[pi[5] + c];
        const int a0 = bt.comp_i[ci * 4], nl = bt.comp_i[ci * 4 + 1], na = nl + bt.comp_i[ci * 4 + 2];
        const int cls = bt.comp_class[ci];
        const int skip = cls >= 0 ? bt.comp_ilplane[ci] : 0;       // planes of a tabulated layer are not listed
        const int np = bt.comp_nplanes[ci] - skip;
        if (threadIdx.x == 0) {
            ps->comp_first[c] = first_plane;
            ps->comp_count[c] = np;
            ps->comp_a0[c] = first_atom + (cls >= 0 ? nl : 0);
            ps->comp_tab[c] = cls >= 0 ? (long long)cls * sumX : -1LL;
            ps->comp_d001[c] = bt.comp_d[ci * 8 + 0];
            ps->comp_delta[c] = bt.comp_d[ci * 8 + 2];
        }
        for (int a = threadIdx.x; a < na; a += blockDim.x) {
            const int l = first_atom + a;
            const int pl = bt.atom_plane[a0 + a] - skip;
            const int t = bt.atom_type[a0 + a];         // None atom type -> contributes 0 (atoms.py:65-67)
            ps->atom_po[l] = make_double2(t >= 0 ? bt.atom_d[(a0 + a) * 2] : 0.0, __longlong_as_double(t >= 0 ? t * sumX : 0LL));
            if (pl < 0) continue;
            if (a == 0 || bt.atom_plane[a0 + a - 1] - skip != pl)
                ps->plane_arg[first_plane + pl] = __dmul_rn(PX_TWO_PI, bt.atom_z[a0 + a]);     // 2 * pi * atom.z
            if (a == na - 1 || bt.atom_plane[a0 + a + 1] - skip != pl) ps->plane_end[first_plane + pl] = l + 1;
        }
        first_atom += na;
        first_plane += np;
    }
    if (threadIdx.x == 0) {
        ps->n_atoms = first_atom;
        lpf_setup(ps, bt.phase_d[pb * 8], derived_s);
    }
    if constexpr (DISTINCT) {
        // plane_arg / plane_end written above are replaced by the distinct-plane tables of px_planes_phase (the writes
        // above and below touch the same addresses from different threads: order them)
        __syncthreads();
        const int nu = bt.ph_nu[pb];
        const double *uz = bt.ph_uz + (long long)pb * bt.acap;
        const int *uend = bt.ph_uend + (long long)pb * bt.acap, *useg = bt.ph_useg + (long long)pb * bt.acap;
        for (int u = threadIdx.x; u < nu; u += blockDim.x) {
            ps->plane_arg[u] = __dmul_rn(PX_TWO_PI, uz[u]);
            ps->plane_end[u] = uend[u];
        }
        for (int i = threadIdx.x; i < first_plane; i += blockDim.x) ps->useg[i] = useg[i];
        if (threadIdx.x == 0) ps->n_u = nu;
    }
}

// one plane: f * exp(i * ((2 pi z) * stl)), f = sum of asf * pn over its atoms     (atoms.py:64)
__device__ __forceinline__ void plane_term(double f, double arg2piz, double stl, double &sr, double &si) {
    double sn, cs;
    px_sincos(__dmul_rn(arg2piz, stl), sn, cs);
    sr = fma(f, cs, sr);
    si = fma(f, sn, si);
}

// phi = exp(2 pi stl (i d001 - pi delta_c stl))      (components.py:53)
__device__ __forceinline__ void phase_factor(double stl, double d001, double delta_c, double &pr, double &pi_) {
    const double A = __dmul_rn(PX_TWO_PI, stl);
    double sn, cs;
    px_sincos(__dmul_rn(A, d001), sn, cs);
    if (delta_c != 0.0) {                                   // block-uniform branch
        const double arg[1] = {__dmul_rn(A, -__dmul_rn(__dmul_rn(PX_PI, delta_c), stl))};
        double e[1];
        px_exp_neg_n<1>(arg, e);
        pr = e[0] * cs;
        pi_ = e[0] * sn;
    } else {                                                // exp(-0.0) == 1 exactly
        pr = cs;
        pi_ = sn;
    }
}

// structure factor SF_c and phase factor phi_c of component c at one point
// (atoms.py:40-67, components.py:18-54)
__device__ __forceinline__ void component_factors(const PhaseSmem *ps, int c, double stl, const double *__restrict__ asf_col,
                                                  const double2 *__restrict__ tab_col, long long sumX, double &ur, double &ui,
                                                  double &pr, double &pi_) {
    double sr = 0.0, si = 0.0;
    if (ps->comp_tab[c] >= 0) {                          // block-uniform: the tabulated layer stack
        const double2 t = __ldg(tab_col + ps->comp_tab[c]);
        sr = t.x;
        si = t.y;
    }
    const int p0 = ps->comp_first[c], p1 = p0 + ps->comp_count[c];
    int a = ps->comp_a0[c];
    for (int p = p0; p < p1; ++p) {
        const int end = ps->plane_end[p];
        double f = 0.0;
        for (; a < end; ++a) {
            const double2 po = ps->atom_po[a];
            f = fma(__ldg(asf_col + __double_as_longlong(po.y)), po.x, f);
        }
        plane_term(f, ps->plane_arg[p], stl, sr, si);
    }
    ur = sr;
    ui = si;
    phase_factor(stl, ps->comp_d001[c], ps->comp_delta[c], pr, pi_);
}

// sin and cos of N independent arguments in lock step (every constant is fetched once for N FMAs, the N dependent chains
// interleave: DFMA latency is 8.5 cycles, tools/microbench.cu), table-assisted: a = k h + r, h = pi / 64,
// (S, C) = (sin, cos)(k h) from the 128-entry table PX_SCT (tools/make_sincos_table.py),
//   sin a = S + (S (cos r - 1) + C sin r),  cos a = C + (C (cos r - 1) - S sin r),  |r| <= pi / 128:
// degree-7 / degree-6 Taylor polynomials are exact to 1e-20 / 4e-18 there.  17 FP64 + 4 other instructions per argument;
// the quadrant reduction with degree-13 / 14 minimax polynomials it replaces took 21 + 13 (quadrant selection, sign
// flips).  Worst absolute error 1.6e-16 on |a| <= 200 (the generator's report; the polynomial version: 1.7e-16).
// SMEM: the table is the caller's copy in shared memory (k_phase_small), else the global one through the read-only path.
// px_sincos (scalar) performs the same operations in the same order: every kernel gets bit-identical values.
template <int N, bool SMEM = false>
__device__ __forceinline__ void px_sincos_n(const double (&a)[N], double (&s)[N], double (&c)[N], const double2 *__restrict__ smem_tab = nullptr) {
    double q[N], r[N];
    int n[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double t = fma(a[k], PX_ST[0], PX_ST[1]);
        q[k] = t - PX_ST[1];
        n[k] = __double2loint(t) & (PX_SCT_N - 1);
    }
    double2 sc[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
        if constexpr (SMEM) sc[k] = smem_tab[n[k]];
        else sc[k] = __ldg(reinterpret_cast<const double2 *>(PX_SCT) + n[k]);
    }
#pragma unroll
    for (int k = 0; k < N; ++k) r[k] = fma(q[k], PX_ST[2], a[k]);
#pragma unroll
    for (int k = 0; k < N; ++k) r[k] = fma(q[k], PX_ST[3], r[k]);
    double r2[N], ps[N], pc[N];
#pragma unroll
    for (int k = 0; k < N; ++k) { r2[k] = r[k] * r[k]; ps[k] = fma(r2[k], PX_ST[4], PX_ST[5]); pc[k] = fma(r2[k], PX_ST[7], PX_ST[8]); }
#pragma unroll
    for (int k = 0; k < N; ++k) { ps[k] = fma(ps[k], r2[k], PX_ST[6]); pc[k] = fma(pc[k], r2[k], PX_ST[9]); }
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double sr = fma(r[k] * r2[k], ps[k], r[k]), cm1 = r2[k] * pc[k];
        s[k] = fma(sc[k].y, sr, fma(sc[k].x, cm1, sc[k].x));
        c[k] = fma(-sc[k].x, sr, fma(sc[k].y, cm1, sc[k].y));
    }
}
template <int N, bool TAB>
__device__ __forceinline__ void px_sincos_sel(const double2 *__restrict__ tab, const double (&a)[N], double (&s)[N], double (&c)[N]) {
    px_sincos_n<N, TAB>(a, s, c, tab);
}

// structure factor and phase factor of component c at the N points gx0 + k * stride of one thread
// (same arithmetic per point as component_factors)
template <int N, bool TAB = false>
__device__ __forceinline__ void component_factors_n(const PhaseSmem *ps, int c, const double (&stl)[N],
                                                    const double *__restrict__ asf_col, const double2 *__restrict__ tab_col,
                                                    int stride, double (&ur)[N], double (&ui)[N], double (&pr)[N],
                                                    double (&pim)[N], const double2 *__restrict__ sct = nullptr) {
    double sr[N], si[N];
    if (ps->comp_tab[c] >= 0) {                          // block-uniform: the tabulated layer stack
        const double2 *tc = tab_col + ps->comp_tab[c];
#pragma unroll
        for (int k = 0; k < N; ++k) { const double2 t = __ldg(tc + k * stride); sr[k] = t.x; si[k] = t.y; }
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) { sr[k] = 0.0; si[k] = 0.0; }
    }
    const int p0 = ps->comp_first[c], p1 = p0 + ps->comp_count[c];
    int a = ps->comp_a0[c];
    for (int p = p0; p < p1; ++p) {
        const int end = ps->plane_end[p];
        double f[N];
#pragma unroll
        for (int k = 0; k < N; ++k) f[k] = 0.0;
        for (; a < end; ++a) {
            const double2 po = ps->atom_po[a];
            const double *col = asf_col + __double_as_longlong(po.y);
#pragma unroll
            for (int k = 0; k < N; ++k) f[k] = fma(__ldg(col + k * stride), po.x, f[k]);
        }
        const double arg = ps->plane_arg[p];
        double ang[N], sn[N], cs[N];
#pragma unroll
        for (int k = 0; k < N; ++k) ang[k] = __dmul_rn(arg, stl[k]);
        px_sincos_sel<N, TAB>(sct, ang, sn, cs);
#pragma unroll
        for (int k = 0; k < N; ++k) { sr[k] = fma(f[k], cs[k], sr[k]); si[k] = fma(f[k], sn[k], si[k]); }
    }
    const double d001 = ps->comp_d001[c], delta_c = ps->comp_delta[c];
    double ang[N], sn[N], cs[N];
#pragma unroll
    for (int k = 0; k < N; ++k) ang[k] = __dmul_rn(__dmul_rn(PX_TWO_PI, stl[k]), d001);
    px_sincos_sel<N, TAB>(sct, ang, sn, cs);
#pragma unroll
    for (int k = 0; k < N; ++k) { ur[k] = sr[k]; ui[k] = si[k]; }
    if (delta_c != 0.0) {                                   // block-uniform branch
        double arg[N], e[N];
#pragma unroll
        for (int k = 0; k < N; ++k) arg[k] = __dmul_rn(__dmul_rn(PX_TWO_PI, stl[k]), -__dmul_rn(__dmul_rn(PX_PI, delta_c), stl[k]));
        px_exp_neg_n<N>(arg, e);
#pragma unroll
        for (int k = 0; k < N; ++k) { pr[k] = e[k] * cs[k]; pim[k] = e[k] * sn[k]; }
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) { pr[k] = cs[k]; pim[k] = sn[k]; }
    }
}

// structure factors and phase factors of ALL G components at the N points of one thread, one complex exponential per
// DISTINCT plane of the phase (load_phase_lists<true>); per plane the arithmetic is that of component_factors_n
template <int G, int N, bool TAB = false>
__device__ __forceinline__ void phase_factors_n(const PhaseSmem *ps, const double (&stl)[N], const double *__restrict__ asf_col,
                                                const double2 *__restrict__ tab_col, int stride, double (&ur)[G][N],
                                                double (&ui)[G][N], double (&pr)[G][N], double (&pim)[G][N],
                                                const double2 *__restrict__ sct = nullptr) {
#pragma unroll
    for (int c = 0; c < G; ++c) {
        if (ps->comp_tab[c] >= 0) {                      // block-uniform: the tabulated layer stack of the component
            const double2 *tc = tab_col + ps->comp_tab[c];
#pragma unroll
            for (int k = 0; k < N; ++k) { const double2 t = __ldg(tc + k * stride); ur[c][k] = t.x; ui[c][k] = t.y; }
        } else {
#pragma unroll
            for (int k = 0; k < N; ++k) { ur[c][k] = 0.0; ui[c][k] = 0.0; }
        }
    }
    const int nu = ps->n_u;
    int sg = 0;
    for (int u = 0; u < nu; ++u) {
        const double arg = ps->plane_arg[u];
        double ang[N], sn[N], cs[N];
#pragma unroll
        for (int k = 0; k < N; ++k) ang[k] = __dmul_rn(arg, stl[k]);
        px_sincos_sel<N, TAB>(sct, ang, sn, cs);
        const int sg_end = ps->plane_end[u];
        for (; sg < sg_end; ++sg) {
            const int info = ps->useg[sg];
            const int a1 = (info >> 9) & 0x1ff, comp = info >> 18;
            double f[N];
#pragma unroll
            for (int k = 0; k < N; ++k) f[k] = 0.0;
            for (int a = info & 0x1ff; a < a1; ++a) {
                const double2 po = ps->atom_po[a];
                const double *col = asf_col + __double_as_longlong(po.y);
#pragma unroll
                for (int k = 0; k < N; ++k) f[k] = fma(__ldg(col + k * stride), po.x, f[k]);
            }
#pragma unroll
            for (int c = 0; c < G; ++c)
                if (comp == c) {                            // block-uniform
#pragma unroll
                    for (int k = 0; k < N; ++k) { ur[c][k] = fma(f[k], cs[k], ur[c][k]); ui[c][k] = fma(f[k], sn[k], ui[c][k]); }
                }
        }
    }
#pragma unroll
    for (int c = 0; c < G; ++c) {
        const double d001 = ps->comp_d001[c], delta_c = ps->comp_delta[c];
        double ang[N], sn[N], cs[N];
#pragma unroll
        for (int k = 0; k < N; ++k) ang[k] = __dmul_rn(__dmul_rn(PX_TWO_PI, stl[k]), d001);
        px_sincos_sel<N, TAB>(sct, ang, sn, cs);
        if (delta_c != 0.0) {                               // block-uniform branch
            double arg[N], e[N];
#pragma unroll
            for (int k = 0; k < N; ++k) arg[k] = __dmul_rn(__dmul_rn(PX_TWO_PI, stl[k]), -__dmul_rn(__dmul_rn(PX_PI, delta_c), stl[k]));
            px_exp_neg_n<N>(arg, e);
#pragma unroll
            for (int k = 0; k < N; ++k) { pr[c][k] = e[k] * cs[k]; pim[c][k] = e[k] * sn[k]; }
        } else {
#pragma unroll
            for (int k = 0; k < N; ++k) { pr[c][k] = cs[k]; pim[c][k] = sn[k]; }
        }
    }
}

// everything after Re(a^T S b): absolute scale, LPF, machine correction (phases.py:158,81-95, specimen.py:57-62)
__device__ __forceinline__ double finish_point(double tr, double abs_scale, int flags, const PhaseSmem *ps,
                                               const StaticDev &st, long long gx) {
    double I = tr * abs_scale;
    if (flags & 2) I = I * lpf_point(ps, st.sin_t[gx], st.inv_sin[gx], st.c2t2[gx]);
    if (flags & 4) I = I * st.corr[gx];
    return I;
}

// invalid probabilities (decided on the device for model-driven phases) -> zeros (phases.py:109-111)
__device__ __forceinline__ bool phase_is_invalid(const BatchDev &bt, int pb, int job, int X, int x0, int tile) {
    if (bt.phase_valid[pb]) return false;
    double *out = bt.intensities + bt.job_out[job];
    for (int x = x0 + threadIdx.x; x < X && x < x0 + tile; x += blockDim.x) out[x] = 0.0;
    return true;
}

// ---------------------------------------------------------------------------------
// phase images: warp per phase, once per generation (after the prologue kernels).  The block runs the staging code the
// phase kernels run per block (load_phase_lists: job -> phase -> component list -> components -> atoms, a chain of six
// dependent global loads; the Lorentz-polarisation constants with three divisions; WG from W) ONCE and lays the result
// out contiguously (BatchDev::phase_img).  A block of k_phase_small then fills its shared memory with one round of
// independent loads.  ncu on c2 before this: 20 % of the warps' time went into the staging code of a block (long
// scoreboard on the chain, math-pipe throttle on the divisions) for 4 % of the executed instructions.
// ---------------------------------------------------------------------------------
// The image is written in two halves, each by the warp at the end of the prologue chain it depends on (the chains run side
// by side on two streams):
//   k_phase_prologue       mean / absolute scale / ntop / flags, P, WG, coefficients (px_image_coef)
//   k_phase_planes_image   z-stretch of the phase's components -> distinct planes -> lists, components,
//                          Lorentz-polarisation constants
// k_phase_planes_image is the whole second chain in one launch, warp per phase: lane c stretches component c of the phase
// (a component that several phases share is stretched by each of them: the same values to the same places), then the warp
// takes the distinct planes and stages the lists exactly as a phase-kernel block would.  Three launches, each a latency
// chain of its own, became one (c2: 9 + 8 + 12 us under ncu -> see DESIGN.md).
union PlanesImageSmem {
    PlanesSmem planes;
    PhaseSmem phase;
};
// thread per component: the z-stretch as a launch of its own, for batches whose phases share components (each phase
// stretching them again inside k_phase_planes_image costs more than the launch: c4, six phases over shared components)
__global__ void k_component_z(BatchDev bt) {
    const int ci = blockIdx.x * blockDim.x + threadIdx.x;
    if (ci < bt.NC) px_component_z(bt, ci);
}
__global__ void __launch_bounds__(32) k_phase_planes_image(StaticDev st, BatchDev bt, int stretch) {
    __shared__ PlanesImageSmem sm;
    const int pb = blockIdx.x, lane = threadIdx.x;
    const int *pi = bt.phase_i + pb * 8;
    const int G = pi[0];
    if (stretch && (pi[4] & 1) && !(pi[4] & 8))
        for (int c = lane; c < G; c += 32) px_component_z(bt, bt.phase_comp[pi[5] + c]);
    __syncwarp();
    px_planes_phase(bt, pb, lane, &sm.planes);
    __syncwarp();
    if (!(pi[4] & 1) || (pi[4] & 8)) return;                 // no components / raw pattern: no phase kernel reads an image
    double *img = bt.phase_img + (long long)pb * bt.img_stride;
    const int AE = bt.img_ae;
    double *pa = img + PX_IMG_HDR;
    PhaseSmem &ps = sm.phase;
    if (G > 1) load_phase_lists<true>(bt, pb, G, &ps, st.sumX, st.derived);
    else load_phase_lists<false>(bt, pb, G, &ps, st.sumX, st.derived);
    if (lane < st.S && lane < 8) lpf_consts(bt.phase_d[pb * 8], st.derived + lane * 8, ps.lpf_all[lane]);
    __syncwarp();
    for (int i = 4 + lane; i < PX_IMG_HDR; i += 32) img[i] = reinterpret_cast<const double *>(&ps)[i];
    double2 *po = reinterpret_cast<double2 *>(pa + AE);
    int *pe = reinterpret_cast<int *>(pa + 3 * AE), *us = pe + AE;
    const int na = ps.n_atoms;
    for (int i = lane; i < AE; i += 32) {                    // (entries past the phase's counts are never read: zeros)
        const bool in = i < na;
        pa[i] = in ? ps.plane_arg[i] : 0.0;
        po[i] = in ? ps.atom_po[i] : make_double2(0.0, 0.0);
        pe[i] = in ? ps.plane_end[i] : 0;
        us[i] = (in && G > 1) ? ps.useg[i] : 0;
    }
}

// helpers of the in-place structured series (k_phase_small, STRUCT = 1): states are digit strings (i1 .. iD) in base
// G, R = G^D; px_rotl moves the leading digit to the end, px_rot applies it s times
template <int V> struct px_int { static constexpr int value = V; };
__host__ __device__ constexpr int px_digits(int R, int G) { int d = 0; for (int v = 1; v < R; v *= G) ++d; return d; }
__host__ __device__ constexpr int px_rot(int L, int s, int R, int G) {
    for (int i = 0; i < s; ++i) L = (L % (R / G)) * G + L / (R / G);
    return L;
}

// ---- ranks up to 9 (structured transition matrices: up to 27): every vector in registers, PPT points per thread ----
// A thread owns the points x0 + t + k * TPB, k < PPT: the warp-uniform work (shared-memory
// reads of the atom / plane lists and of P, loop control, coefficient fetches) is paid once for
// PPT points and the dependent FMA chains of the PPT points interleave.  Per-point tables are
// padded by PX_TABLE_PAD doubles, so the loads of points beyond the end of a pattern are legal;
// only the stores are guarded.
// STRUCT (the host checks the matrix entry by entry, or knows it from the probability model):
//  1  P has the structure every Reichweite >= 2 model produces (base_models.py:243-261): state I = (i1 .. iR) only
//     reaches the states (i2 .. iR, j), i.e. row I is zero outside the columns (I mod R/G) * G + j, j < G: the product
//     y = P w takes G instead of R multiply-adds per state;
//  2  every row of P is the same vector w (Reichweite 0, R0models.py:80-83): Q^n = 1 (w o phi)^T q^(n-1) with the
//     complex SCALAR q = sum_J w_J phi_J, so the series is a scalar Horner recurrence whatever G is:
//     sum_n c_n Q^n b = 1 * (sum_J w_J phi_J b_J) * sum_n c_n q^(n-1).
//
// A block walks TILES consecutive tiles of TPB * PPT points of its pattern: staging the phase's parameter block (lists,
// matrices, coefficients, LPF constants: a chain of dependent global loads) is paid once per TILES tiles.
template <int R, int G, int PPT, int TPB, int MINB, int STRUCT = 0, int TILES = 1>
__global__ void __launch_bounds__(TPB, MINB) k_phase_small(StaticDev st, BatchDev bt, const int *__restrict__ jobs) {
    constexpr int REPS = R / G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PhaseSmem *ps = reinterpret_cast<PhaseSmem *>(smem_raw);
    double *sP = reinterpret_cast<double *>(smem_raw + sizeof(PhaseSmem));   // [R][R]
    double *sWG = sP + R * R;                                               // [G][R]
    double *sCoef = sWG + G * R;                                            // [ntop+1]
    // the sine / cosine table: a copy in shared memory for the instantiations that are at their register limit anyway
    // (MINB > 0: the c2 kernel, 16 blocks per SM), the global table through the read-only path for the others -- their
    // occupancy is limited by shared memory, and 2 KB more per block would cost blocks (c1: 25 -> 20 per SM)
    constexpr bool SCT_SMEM = PX_SCT_SMEM && MINB > 0;
    double2 *sSct = reinterpret_cast<double2 *>(smem_raw + bt.img_sct_off);   // [PX_SCT_N] (sin, cos)(k pi / 64), px_sincos_n

    const int job = jobs[blockIdx.x];
    const int pb = bt.job_pb[job], s = bt.job_spec[job];
    const int X = st.spec_npts[s];
    const int xblock = blockIdx.y * (TPB * PPT * TILES);
    if (xblock >= X || phase_is_invalid(bt, pb, job, X, xblock, TPB * PPT * TILES)) return;
    constexpr bool DISTINCT = G > 1;        // one component: nothing to share between components
    {
        // staging from the phase image (k_phase_prologue + k_phase_planes_image): one round of independent loads
        const double *img = bt.phase_img + (long long)pb * bt.img_stride;
        const int AE = bt.img_ae;
        double *hdr = reinterpret_cast<double *>(ps);
        for (int i = threadIdx.x; i < PX_IMG_HDR; i += TPB) hdr[i] = img[i];
        if (threadIdx.x < 4) (&ps->lpf_q)[threadIdx.x] = img[4 + 4 * s + threadIdx.x];       // lpf_all[s]
        const double *pa = img + PX_IMG_HDR;
        const double2 *po = reinterpret_cast<const double2 *>(pa + AE);
        const int *pe = reinterpret_cast<const int *>(pa + 3 * AE), *us = pe + AE;
        for (int i = threadIdx.x; i < AE; i += TPB) {
            ps->plane_arg[i] = pa[i];
            ps->atom_po[i] = po[i];
            ps->plane_end[i] = pe[i];
            if constexpr (DISTINCT) ps->useg[i] = us[i];
        }
        const double *mat = pa + 4 * AE;
        for (int i = threadIdx.x; i < R * R; i += TPB) sP[i] = mat[i];
        for (int i = threadIdx.x; i < G * R; i += TPB) sWG[i] = mat[R * R + i];
        const double *cf = mat + bt.img_mcap;
        for (int i = threadIdx.x; i < bt.img_ncopy; i += TPB) sCoef[i] = cf[i];
        if constexpr (SCT_SMEM)
            for (int i = threadIdx.x; i < PX_SCT_N; i += TPB) sSct[i] = reinterpret_cast<const double2 *>(PX_SCT)[i];
    }
    __syncthreads();
    const int ntop = (int)ps->px[2];
    const double mean = ps->px[0], abs_scale = ps->px[1];
    const int flags = (int)ps->px[3];

#pragma unroll 1
    for (int x0 = xblock; x0 < xblock + TPB * PPT * TILES; x0 += TPB * PPT) {
    const int x = x0 + threadIdx.x;
    if (x >= X) break;
    const long long gx = st.spec_off[s] + x;
    double stl[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) stl[k] = st.stl[gx + k * TPB];

    double ur[G][PPT], ui[G][PPT], pr[G][PPT], pim[G][PPT], gr[G][PPT], gi[G][PPT];
    if constexpr (DISTINCT) phase_factors_n<G, PPT, SCT_SMEM>(ps, stl, st.asf + gx, bt.layer_tab + gx, TPB, ur, ui, pr, pim, sSct);
#pragma unroll
    for (int c = 0; c < G; ++c) {
        if constexpr (!DISTINCT) component_factors_n<PPT, SCT_SMEM>(ps, c, stl, st.asf + gx, bt.layer_tab + gx, TPB, ur[c], ui[c], pr[c], pim[c], sSct);
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            gr[c][k] = pr[c][k] * ur[c][k] + pim[c][k] * ui[c][k];          // g = phi * conj(u)
            gi[c][k] = pim[c][k] * ur[c][k] - pr[c][k] * ui[c][k];
        }
    }
    double yr[R][PPT], yi[R][PPT];
    int y_layout = 0;         // STRUCT = 1: the registers hold the state vector rotated by this many digits
#pragma unroll
    for (int I = 0; I < R; ++I)
#pragma unroll
        for (int k = 0; k < PPT; ++k) { yr[I][k] = 0.0; yi[I][k] = 0.0; }
    if constexpr (STRUCT == 2) {
        constexpr int REPS2 = R / G;
        double qr[PPT], qi[PPT], br[PPT], bi[PPT], sr[PPT], si[PPT];
#pragma unroll
        for (int k = 0; k < PPT; ++k) { qr[k] = 0.0; qi[k] = 0.0; br[k] = 0.0; bi[k] = 0.0; sr[k] = 0.0; si[k] = 0.0; }
#pragma unroll
        for (int J = 0; J < R; ++J) {
            const double w = sP[J];                     // first row = every row
            const int c = J / REPS2;
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                qr[k] = fma(w, pr[c][k], qr[k]); qi[k] = fma(w, pim[c][k], qi[k]);
                br[k] = fma(w, gr[c][k], br[k]); bi[k] = fma(w, gi[c][k], bi[k]);
            }
        }
        for (int n = ntop; n >= 1; --n) {               // sigma = c_1 + q (c_2 + q (c_3 + ...))
            const double cn = sCoef[n];
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                const double nr = fma(sr[k], qr[k], fma(-si[k], qi[k], cn));
                si[k] = fma(sr[k], qi[k], si[k] * qr[k]);
                sr[k] = nr;
            }
        }
#pragma unroll
        for (int I = 0; I < R; ++I)
#pragma unroll
            for (int k = 0; k < PPT; ++k) {             // y_I = beta sigma
                yr[I][k] = br[k] * sr[k] - bi[k] * si[k];
                yi[I][k] = br[k] * si[k] + bi[k] * sr[k];
            }
    } else if constexpr (R == 1) {
        // scalar series: y = (sum_n c_n q^n) b, q = P00 phi, by Horner on the real coefficients (4 flop-pairs / term)
        const double p00 = sP[0];
        double qr[PPT], qi[PPT], sr[PPT], si[PPT];
#pragma unroll
        for (int k = 0; k < PPT; ++k) { qr[k] = p00 * pr[0][k]; qi[k] = p00 * pim[0][k]; sr[k] = 0.0; si[k] = 0.0; }
        for (int n = ntop; n >= 1; --n) {
            const double cn = sCoef[n];
#pragma unroll
            for (int k = 0; k < PPT; ++k) {             // s <- (s + c_n) q
                const double tr_ = sr[k] + cn;
                const double nr = fma(tr_, qr[k], -si[k] * qi[k]);
                si[k] = fma(tr_, qi[k], si[k] * qr[k]);
                sr[k] = nr;
            }
        }
#pragma unroll
        for (int k = 0; k < PPT; ++k) {                 // y = s conj(u)
            yr[0][k] = sr[k] * ur[0][k] + si[k] * ui[0][k];
            yi[0][k] = si[k] * ur[0][k] - sr[k] * ui[0][k];
        }
    } else if constexpr (R == 2) {
        // 2 x 2: Cayley-Hamilton Q^2 = t Q - d Id (t = trace, d = determinant) makes the powers a
        // three-term recurrence, summed by Clenshaw's backward recurrence B_k = c_k + t B_{k+1} - d B_{k+2}
        // on complex SCALARS (8 flop-pairs / term instead of 20 for the vector Horner step):
        //   sum_{n>=1} c_n Q^n = B_1 Q - d B_2 Id.
        // |eigenvalues of Q| <= 1 (P stochastic, |phi| <= 1), so the recurrence does not amplify
        // rounding errors; measured against an 80-bit evaluation it is as accurate as the Horner
        // form (7e-15 relative on c2).
        constexpr int C1 = 1 / REPS;                    // component of state 1 (0 when G == 1)
        const double p00 = sP[0], p01 = sP[1], p10 = sP[2], p11 = sP[3];
        double tr_[PPT], ti_[PPT], dr[PPT], di[PPT], b1r[PPT], b1i[PPT], b2r[PPT], b2i[PPT];
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            // Q[I][J] = P[I][J] phi_J
            tr_[k] = fma(p00, pr[0][k], p11 * pr[C1][k]);
            ti_[k] = fma(p00, pim[0][k], p11 * pim[C1][k]);
            const double ffr = pr[0][k] * pr[C1][k] - pim[0][k] * pim[C1][k];   // phi_0 phi_1
            const double ffi = pr[0][k] * pim[C1][k] + pim[0][k] * pr[C1][k];
            const double dp = p00 * p11 - p01 * p10;
            dr[k] = dp * ffr;
            di[k] = dp * ffi;
            b1r[k] = 0.0; b1i[k] = 0.0; b2r[k] = 0.0; b2i[k] = 0.0;
        }
        for (int n = ntop; n >= 1; --n) {
            const double cn = sCoef[n];
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                const double nr = fma(tr_[k], b1r[k], fma(-ti_[k], b1i[k], fma(-dr[k], b2r[k], fma(di[k], b2i[k], cn))));
                const double ni = fma(tr_[k], b1i[k], fma(ti_[k], b1r[k], fma(-dr[k], b2i[k], -di[k] * b2r[k])));
                b2r[k] = b1r[k]; b2i[k] = b1i[k];
                b1r[k] = nr; b1i[k] = ni;
            }
        }
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            // y = B_1 (Q b) - (d B_2) b, b_J = conj(u_comp(J)), (Q b)_I = sum_J P[I][J] g_J, g = phi conj(u)
            const double e_r = dr[k] * b2r[k] - di[k] * b2i[k], e_i = dr[k] * b2i[k] + di[k] * b2r[k];
            const double q0r = fma(p00, gr[0][k], p01 * gr[C1][k]), q0i = fma(p00, gi[0][k], p01 * gi[C1][k]);
            const double q1r = fma(p10, gr[0][k], p11 * gr[C1][k]), q1i = fma(p10, gi[0][k], p11 * gi[C1][k]);
            yr[0][k] = b1r[k] * q0r - b1i[k] * q0i - (e_r * ur[0][k] + e_i * ui[0][k]);
            yi[0][k] = b1r[k] * q0i + b1i[k] * q0r - (e_i * ur[0][k] - e_r * ui[0][k]);
            yr[1][k] = b1r[k] * q1r - b1i[k] * q1i - (e_r * ur[C1][k] + e_i * ui[C1][k]);
            yi[1][k] = b1r[k] * q1i + b1i[k] * q1r - (e_i * ur[C1][k] - e_r * ui[C1][k]);
        }
    } else if constexpr (STRUCT == 1) {
        // Structured transitions, IN PLACE.  The new value of state (a, i2 .. iD) is a combination of the G old values of
        // the states (i2 .. iD, j), and those G old values are needed by exactly the G new states a = 0 .. G-1: a group of
        // G registers is read, then overwritten with the G results.  The result for (a, g) lands where (g, a) was, i.e.
        // after a step the registers hold the vector rotated by one digit; after D steps the layout is back, so the
        // loop body is D steps with compile-time layouts.  One vector instead of two: half the registers of the y / w
        // version.
        constexpr int T = R / G, D = px_digits(R, G);
        auto substep = [&](auto S, const double cn) {
            constexpr int sft = decltype(S)::value;
            double cgr[G][PPT], cgi[G][PPT];
#pragma unroll
            for (int c = 0; c < G; ++c)
#pragma unroll
                for (int k = 0; k < PPT; ++k) { cgr[c][k] = cn * gr[c][k]; cgi[c][k] = cn * gi[c][k]; }
#pragma unroll
            for (int g = 0; g < T; ++g) {
                double wr[G][PPT], wi[G][PPT];
#pragma unroll
                for (int j = 0; j < G; ++j) {                 // w_J = phi_c(J) y_J + c_n g_c(J), J = (g, j)
                    const int J = g * G + j, c = J / REPS, pj = px_rot(J, sft, R, G);
#pragma unroll
                    for (int k = 0; k < PPT; ++k) {
                        wr[j][k] = fma(pr[c][k], yr[pj][k], fma(-pim[c][k], yi[pj][k], cgr[c][k]));
                        wi[j][k] = fma(pr[c][k], yi[pj][k], fma(pim[c][k], yr[pj][k], cgi[c][k]));
                    }
                }
#pragma unroll
                for (int a = 0; a < G; ++a) {                 // y_(a, g) = sum_j P[(a, g)][(g, j)] w_(g, j)
                    const int I = a * T + g, pa = px_rot(g * G + a, sft, R, G);
                    double ar[PPT], ai[PPT];
#pragma unroll
                    for (int k = 0; k < PPT; ++k) { ar[k] = 0.0; ai[k] = 0.0; }
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        const double p = sP[I * R + g * G + j];
#pragma unroll
                        for (int k = 0; k < PPT; ++k) { ar[k] = fma(p, wr[j][k], ar[k]); ai[k] = fma(p, wi[j][k], ai[k]); }
                    }
#pragma unroll
                    for (int k = 0; k < PPT; ++k) { yr[pa][k] = ar[k]; yi[pa][k] = ai[k]; }
                }
            }
        };
        int n = ntop, layout = 0;                             // logical state L lives in register px_rot(L, layout)
#define PX_SUBSTEP(S_) if (n >= 1) { substep(px_int<S_>{}, sCoef[n]); --n; layout = (S_ + 1) % D; }
        while (n >= 1) {
            PX_SUBSTEP(0)
            if constexpr (D > 1) PX_SUBSTEP(1)
            if constexpr (D > 2) PX_SUBSTEP(2)
            if constexpr (D > 3) PX_SUBSTEP(3)
        }
#undef PX_SUBSTEP
        static_assert(D <= 4, "Reichweite of the structured in-place series");
        y_layout = layout;                                    // the trace below reads the rotated registers
    } else {
        for (int n = ntop; n >= 1; --n) {
            const double cn = sCoef[n];
            double wr[R][PPT], wi[R][PPT];
    #pragma unroll
            for (int I = 0; I < R; ++I) {                     // w = phi o (y + c_n b)
                const int c = I / REPS;
    #pragma unroll
                for (int k = 0; k < PPT; ++k) {
                    wr[I][k] = fma(pr[c][k], yr[I][k], fma(-pim[c][k], yi[I][k], cn * gr[c][k]));
                    wi[I][k] = fma(pr[c][k], yi[I][k], fma(pim[c][k], yr[I][k], cn * gi[c][k]));
                }
            }
    #pragma unroll
            for (int I = 0; I < R; ++I) {                     // y = P w
                double ar[PPT], ai[PPT];
    #pragma unroll
                for (int k = 0; k < PPT; ++k) { ar[k] = 0.0; ai[k] = 0.0; }
                if constexpr (STRUCT == 1) {
    #pragma unroll
                    for (int j = 0; j < G; ++j) {
                        constexpr int TAILS = R / G;
                        const int J = (I % TAILS) * G + j;      // compile-time: I and j are unrolled
                        const double p = sP[I * R + J];
    #pragma unroll
                        for (int k = 0; k < PPT; ++k) {
                            ar[k] = fma(p, wr[J][k], ar[k]);
                            ai[k] = fma(p, wi[J][k], ai[k]);
                        }
                    }
                } else {
    #pragma unroll
                    for (int J = 0; J < R; ++J) {
                        const double p = sP[I * R + J];
    #pragma unroll
                        for (int k = 0; k < PPT; ++k) {
                            ar[k] = fma(p, wr[J][k], ar[k]);
                            ai[k] = fma(p, wi[J][k], ai[k]);
                        }
                    }
                }
    #pragma unroll
                for (int k = 0; k < PPT; ++k) { yr[I][k] = ar[k]; yi[I][k] = ai[k]; }
            }
        }
    }
    // Re( sum_K a_K (mean b_K + y_K) ), a_K = sum_c u_c WG[c][K], b_K = conj(u_comp(K))
    double tr[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) tr[k] = 0.0;
    auto trace = [&](auto S) {
        constexpr int sft = decltype(S)::value;
#pragma unroll
        for (int K = 0; K < R; ++K) {
            const int ck = K / REPS;
            const int pk = (STRUCT == 1) ? px_rot(K, sft, R, G) : K;       // where y_K lives
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                double ar = 0.0, ai = 0.0;
#pragma unroll
                for (int c = 0; c < G; ++c) {
                    const double w = sWG[c * R + K];
                    ar = fma(ur[c][k], w, ar);
                    ai = fma(ui[c][k], w, ai);
                }
                const double sr = fma(mean, ur[ck][k], yr[pk][k]);
                const double si = fma(-mean, ui[ck][k], yi[pk][k]);
                tr[k] += ar * sr - ai * si;
            }
        }
    };
    if constexpr (STRUCT == 1) {
        if (y_layout == 0) trace(px_int<0>{});
        else if (y_layout == 1) trace(px_int<1>{});
        else if (y_layout == 2) trace(px_int<2>{});
        else trace(px_int<3>{});
    } else {
        trace(px_int<0>{});
    }
    // absolute scale, LPF, machine correction (phases.py:158,81-95, specimen.py:57-62); padded tables: loads past X are safe
    double out[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) out[k] = tr[k] * abs_scale;
    if (flags & 2) {
        double sn[PPT], isn[PPT], c2[PPT], lpf[PPT];
#pragma unroll
        for (int k = 0; k < PPT; ++k) { sn[k] = st.sin_t[gx + k * TPB]; isn[k] = st.inv_sin[gx + k * TPB]; c2[k] = st.c2t2[gx + k * TPB]; }
        lpf_point_n<PPT>(ps, sn, isn, c2, lpf);
#pragma unroll
        for (int k = 0; k < PPT; ++k) out[k] = out[k] * lpf[k];
    }
    if (flags & 4) {
#pragma unroll
        for (int k = 0; k < PPT; ++k) out[k] = out[k] * st.corr[gx + k * TPB];
    }
#pragma unroll
    for (int k = 0; k < PPT; ++k)
        if (x + k * TPB < X) bt.intensities[bt.job_out[job] + x + k * TPB] = out[k];
    }
}

// ---- ranks 16 / 27: w in registers, the new vector staged through shared memory ----
template <int R, int G>
__global__ void __launch_bounds__(PX_TPB) k_phase_staged(StaticDev st, BatchDev bt, const int *__restrict__ jobs) {
    constexpr int REPS = R / G;
    constexpr int RB = (R + 3) / 4;                      // row blocks of 4
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PhaseSmem *ps = reinterpret_cast<PhaseSmem *>(smem_raw);
    double *sPb = reinterpret_cast<double *>(smem_raw + sizeof(PhaseSmem));  // [RB][R][4]: P[4*ib+q][J]
    double *sWG = sPb + RB * R * 4;                                          // [G][R]
    double2 *stage = reinterpret_cast<double2 *>(sWG + G * R + (G * R & 1)); // [R][TPB]
    double *sCoef = reinterpret_cast<double *>(stage + R * PX_TPB);

    const int job = jobs[blockIdx.x];
    const int pb = bt.job_pb[job], s = bt.job_spec[job];
    const int X = st.spec_npts[s];
    const int x0 = blockIdx.y * PX_TPB;
    if (x0 >= X || phase_is_invalid(bt, pb, job, X, x0, PX_TPB)) return;
    const int *pi = bt.phase_i + pb * 8;
    const double *px = bt.phase_x + pb * 4;
    const int ntop = (int)px[2];
    const double mean = px[0], abs_scale = px[1];
    const int flags = pi[4];

    load_phase_lists(bt, pb, G, ps, st.sumX, st.derived + s * 8);
    {
        const double *W = bt.matrices + pi[6];
        const double *P = W + R * R;
        for (int i = threadIdx.x; i < RB * R * 4; i += PX_TPB) {
            const int q = i & 3, J = (i >> 2) % R, ib = (i >> 2) / R;
            const int I = ib * 4 + q;
            sPb[i] = (I < R) ? P[I * R + J] : 0.0;
        }
        for (int i = threadIdx.x; i < G * R; i += PX_TPB) {
            const int c = i / R, K = i % R;
            double acc = 0.0;
            for (int J = c * REPS; J < (c + 1) * REPS; ++J) acc += W[J * R + K];
            sWG[i] = acc;
        }
        const double *coef = bt.coef + (long long)pb * bt.ncap;
        for (int i = threadIdx.x; i <= ntop; i += PX_TPB) sCoef[i] = coef[i];
    }
    __syncthreads();

    const int x = x0 + threadIdx.x;
    if (x >= X) return;
    const long long gx = st.spec_off[s] + x;
    const double stl = st.stl[gx];

    double ur[G], ui[G], pr[G], pim[G], gr[G], gi[G];
#pragma unroll
    for (int c = 0; c < G; ++c) {
        component_factors(ps, c, stl, st.asf + gx, bt.layer_tab + gx, st.sumX, ur[c], ui[c], pr[c], pim[c]);
        gr[c] = pr[c] * ur[c] + pim[c] * ui[c];
        gi[c] = pim[c] * ur[c] - pr[c] * ui[c];
    }
    // w = phi o (y + c_n b) = phi o y + c_n g; for the first step y = 0
    double wr[R], wi[R];
    {
        const double cn = (ntop >= 1) ? sCoef[ntop] : 0.0;
#pragma unroll
        for (int J = 0; J < R; ++J) { wr[J] = cn * gr[J / REPS]; wi[J] = cn * gi[J / REPS]; }
    }
    double2 *mine = stage + threadIdx.x;
    for (int n = ntop; n >= 1; --n) {
        // y = P w, four rows at a time; P[4 ib + q][J] comes as two broadcast LDS.128 per J
        for (int ib = 0; ib < RB; ++ib) {
            const double4 *prow = reinterpret_cast<const double4 *>(sPb + ib * R * 4);
            double a0r = 0, a0i = 0, a1r = 0, a1i = 0, a2r = 0, a2i = 0, a3r = 0, a3i = 0;
#pragma unroll
            for (int J = 0; J < R; ++J) {
                const double4 p = prow[J];
                a0r = fma(p.x, wr[J], a0r); a0i = fma(p.x, wi[J], a0i);
                a1r = fma(p.y, wr[J], a1r); a1i = fma(p.y, wi[J], a1i);
                a2r = fma(p.z, wr[J], a2r); a2i = fma(p.z, wi[J], a2i);
                a3r = fma(p.w, wr[J], a3r); a3i = fma(p.w, wi[J], a3i);
            }
            const int I = ib * 4;
            mine[I * PX_TPB] = make_double2(a0r, a0i);
            if (I + 1 < R) mine[(I + 1) * PX_TPB] = make_double2(a1r, a1i);
            if (I + 2 < R) mine[(I + 2) * PX_TPB] = make_double2(a2r, a2i);
            if (I + 3 < R) mine[(I + 3) * PX_TPB] = make_double2(a3r, a3i);
        }
        if (n > 1) {
            // next w: the component of column J is a compile-time constant here, so phi / g need no selects
            const double cn = sCoef[n - 1];
            double cgr[G], cgi[G];
#pragma unroll
            for (int c = 0; c < G; ++c) { cgr[c] = cn * gr[c]; cgi[c] = cn * gi[c]; }
#pragma unroll
            for (int J = 0; J < R; ++J) {
                const double2 y = mine[J * PX_TPB];
                wr[J] = fma(pr[J / REPS], y.x, fma(-pim[J / REPS], y.y, cgr[J / REPS]));
                wi[J] = fma(pr[J / REPS], y.y, fma(pim[J / REPS], y.x, cgi[J / REPS]));
            }
        }
    }
    // Re( sum_K a_K (mean b_K + y_K) ), a_K = sum_c u_c WG[c][K], b_K = conj(u_comp(K))
    double tr = 0.0;
#pragma unroll
    for (int K = 0; K < R; ++K) {
        double ar = 0.0, ai = 0.0;
#pragma unroll
        for (int c = 0; c < G; ++c) {
            const double w = sWG[c * R + K];
            ar = fma(ur[c], w, ar);
            ai = fma(ui[c], w, ai);
        }
        double2 y = make_double2(0.0, 0.0);
        if (ntop >= 1) y = mine[K * PX_TPB];
        const double sr = fma(mean, ur[K / REPS], y.x);
        const double si = fma(-mean, ui[K / REPS], y.y);
        tr += ar * sr - ai * si;
    }
    bt.intensities[bt.job_out[job] + x] = finish_point(tr, abs_scale, flags, ps, st, gx);
}

// ---- ranks 8..32 on the FP64 tensor cores (DMMA.8x8x4) ----------------------------
// One Horner step y <- P (phi o y + c_n g) for 8 points at a time is the real matrix product
//   Y^T[8 points][r] = W^T[8 points][r] * P^T[r][r]      (re and im as two independent products)
// issued as mma.sync.m8n8k4.f64: A = W^T tile (points x 4 states), B = P^T tile, C = Y^T tile.
// The contraction index may be enumerated in any order, so k-tile (nt, s) is defined to hold
// the states of columns {2q + s, q = 0..3} of n-tile nt: then the lane that owns C[p][2q + s]
// after one step owns exactly the A[p][k = q] element of k-tile (nt, s) of the next step and the
// recurrence needs NO data movement between steps: the accumulator registers, updated in place
// by phi o y + c_n g, are the next A fragments.  P^T fragments stay in registers for the whole
// kernel (KT x NT doubles per lane); the only other FP64 work in the loop is the 6 flop / state
// elementwise update.  tools/microbench.cu: DMMA sustains 37.0 TFLOP/s with two independent
// accumulators per warp (DFMA: 36.2), whereas broadcast-LDS-fed DFMAs (k_phase_staged) stop at
// 25 TFLOP/s for the 2 DFMA / loaded double this product allows.
template <int R>
struct MmaShape {
    static constexpr int NT = (R + 7) / 8;                // n-tiles of 8 states
    static constexpr int REM = R - 8 * (NT - 1);          // states in the last n-tile
    static constexpr bool HALF = REM <= 4;                // ... kept in its even columns: k-tile (NT-1, 1) is all padding
    static constexpr int KT = 2 * NT - (HALF ? 1 : 0);
    // state held by column col of n-tile nt, -1 for padding
    __host__ __device__ static constexpr int state(int nt, int col) {
        if (nt < NT - 1 || !HALF) return (8 * nt + col < R) ? 8 * nt + col : -1;
        return ((col & 1) == 0 && (col >> 1) < REM) ? 8 * nt + (col >> 1) : -1;
    }
};

__device__ __forceinline__ void px_dmma(double &c0, double &c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// v[c] for a per-lane component c known to lie in [cmin, cmax] (compile-time after unrolling:
// a slot whose four lanes share one component costs nothing)
template <int G>
__device__ __forceinline__ double px_pick(const double (&v)[G], int c, int cmin, int cmax) {
    double out = v[cmin];
#pragma unroll
    for (int k = 1; k < G; ++k)
        if (k > cmin && k <= cmax) out = (c >= k) ? v[k] : out;
    return out;
}

template <int R, int G, int MTC>
__global__ void __launch_bounds__(PX_TPB) k_phase_mma(StaticDev st, BatchDev bt, const int *__restrict__ jobs) {
    using SH = MmaShape<R>;
    constexpr int REPS = R / G, NT = SH::NT, KT = SH::KT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PhaseSmem *ps = reinterpret_cast<PhaseSmem *>(smem_raw);
    double *sP = reinterpret_cast<double *>(smem_raw + sizeof(PhaseSmem));   // [R][R]
    double *sWG = sP + R * R;                                               // [G][R]
    double2 *sU = reinterpret_cast<double2 *>(sWG + G * R + ((R * R + G * R) & 1));   // [G][TPB] structure factors
    double2 *sF = sU + G * PX_TPB;                                          // [G][TPB] phase factors
    double *sCoef = reinterpret_cast<double *>(sF + G * PX_TPB);

    const int job = jobs[blockIdx.x];
    const int pb = bt.job_pb[job], s = bt.job_spec[job];
    const int X = st.spec_npts[s];
    const int x0 = blockIdx.y * PX_TPB;
    if (x0 >= X || phase_is_invalid(bt, pb, job, X, x0, PX_TPB)) return;
    const int *pi = bt.phase_i + pb * 8;
    const double *px = bt.phase_x + pb * 4;
    const int ntop = (int)px[2];
    const double mean = px[0], abs_scale = px[1];
    const int flags = pi[4];

    load_phase_lists(bt, pb, G, ps, st.sumX, st.derived + s * 8);
    {
        const double *W = bt.matrices + pi[6];
        const double *P = W + R * R;
        for (int i = threadIdx.x; i < R * R; i += PX_TPB) sP[i] = P[i];
        for (int i = threadIdx.x; i < G * R; i += PX_TPB) {
            const int c = i / R, K = i % R;
            double acc = 0.0;
            for (int J = c * REPS; J < (c + 1) * REPS; ++J) acc += W[J * R + K];
            sWG[i] = acc;
        }
        const double *coef = bt.coef + (long long)pb * bt.ncap;
        for (int i = threadIdx.x; i <= ntop; i += PX_TPB) sCoef[i] = coef[i];
    }
    __syncthreads();

    const int wbase = threadIdx.x & ~31;
    if (x0 + wbase >= X) return;                          // whole warp beyond the pattern
    const int x = x0 + threadIdx.x;
    const bool live = x < X;
    const long long gx = st.spec_off[s] + (live ? x : X - 1);    // idle lanes of the last warp recompute the last point
    const double stl = st.stl[gx];
#pragma unroll
    for (int c = 0; c < G; ++c) {
        double a, b, f, g;
        component_factors(ps, c, stl, st.asf + gx, bt.layer_tab + gx, st.sumX, a, b, f, g);
        sU[c * PX_TPB + threadIdx.x] = make_double2(a, b);
        sF[c * PX_TPB + threadIdx.x] = make_double2(f, g);
    }
    __syncwarp();

    const int lane = threadIdx.x & 31, q = lane & 3, p = lane >> 2;
    // B fragments: lane holds P^T[k = q][n = p] = P[state(nto, p)][state(kt / 2, 2 q + kt % 2)]
    double Pf[KT][NT];
#pragma unroll
    for (int kt = 0; kt < KT; ++kt)
#pragma unroll
        for (int nto = 0; nto < NT; ++nto) {
            const int J = SH::state(kt >> 1, 2 * q + (kt & 1)), I = SH::state(nto, p);
            Pf[kt][nto] = (I >= 0 && J >= 0) ? sP[I * R + J] : 0.0;
        }

    double mytr = 0.0;
#pragma unroll 1
    for (int mg = 0; mg < 4; mg += MTC) {                 // MTC m-tiles (8 points each) in flight
        double fpr[MTC][NT][2], fpi[MTC][NT][2], fgr[MTC][NT][2], fgi[MTC][NT][2];
#pragma unroll
        for (int m = 0; m < MTC; ++m) {
            const int src = wbase + 8 * (mg + m) + p;     // the point of this lane's fragment row
            double tpr[G], tpi[G], tgr[G], tgi[G];
#pragma unroll
            for (int c = 0; c < G; ++c) {
                const double2 u = sU[c * PX_TPB + src], f = sF[c * PX_TPB + src];
                tpr[c] = f.x;
                tpi[c] = f.y;
                tgr[c] = f.x * u.x + f.y * u.y;           // g = phi * conj(u)
                tgi[c] = f.y * u.x - f.x * u.y;
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) {
                    const int I0 = SH::state(nt, sl), I3 = SH::state(nt, 6 + sl), It = SH::state(nt, 2 * q + sl);
                    const int cmin = (I0 >= 0 ? I0 : R - 1) / REPS, cmax = (I3 >= 0 ? I3 : R - 1) / REPS;
                    const int ct = (It >= 0 ? It : R - 1) / REPS;
                    fpr[m][nt][sl] = px_pick<G>(tpr, ct, cmin, cmax);
                    fpi[m][nt][sl] = px_pick<G>(tpi, ct, cmin, cmax);
                    fgr[m][nt][sl] = px_pick<G>(tgr, ct, cmin, cmax);
                    fgi[m][nt][sl] = px_pick<G>(tgi, ct, cmin, cmax);
                }
        }
        double yr[MTC][NT][2], yi[MTC][NT][2];
#pragma unroll
        for (int m = 0; m < MTC; ++m)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) { yr[m][nt][0] = yr[m][nt][1] = 0.0; yi[m][nt][0] = yi[m][nt][1] = 0.0; }
#pragma unroll 1
        for (int n = ntop; n >= 1; --n) {
            const double cn = sCoef[n];
            double wr[MTC][NT][2], wi[MTC][NT][2];
#pragma unroll
            for (int m = 0; m < MTC; ++m)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int sl = 0; sl < 2; ++sl) {      // w = phi o y + c_n g, in the C layout = next A layout
                        wr[m][nt][sl] = fma(fpr[m][nt][sl], yr[m][nt][sl], fma(-fpi[m][nt][sl], yi[m][nt][sl], cn * fgr[m][nt][sl]));
                        wi[m][nt][sl] = fma(fpr[m][nt][sl], yi[m][nt][sl], fma(fpi[m][nt][sl], yr[m][nt][sl], cn * fgi[m][nt][sl]));
                    }
#pragma unroll
            for (int m = 0; m < MTC; ++m)
#pragma unroll
                for (int nto = 0; nto < NT; ++nto) {
                    double r0 = 0.0, r1 = 0.0, i0 = 0.0, i1 = 0.0;
#pragma unroll
                    for (int kt = 0; kt < KT; ++kt) {
                        px_dmma(r0, r1, wr[m][kt >> 1][kt & 1], Pf[kt][nto]);
                        px_dmma(i0, i1, wi[m][kt >> 1][kt & 1], Pf[kt][nto]);
                    }
                    yr[m][nto][0] = r0; yr[m][nto][1] = r1;
                    yi[m][nto][0] = i0; yi[m][nto][1] = i1;
                }
        }
        // Re( sum_K a_K (mean b_K + y_K) ) over this lane's columns, then over the four lanes of the row
#pragma unroll
        for (int m = 0; m < MTC; ++m) {
            const int src = wbase + 8 * (mg + m) + p;
            double tur[G], tui[G];
#pragma unroll
            for (int c = 0; c < G; ++c) {
                const double2 u = sU[c * PX_TPB + src];
                tur[c] = u.x;
                tui[c] = u.y;
            }
            double part = 0.0;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) {
                    if (SH::state(nt, sl) < 0) continue;              // column pair of padding only
                    const int It = SH::state(nt, 2 * q + sl);
                    const int K = It >= 0 ? It : 0;
                    double ar = 0.0, ai = 0.0;
#pragma unroll
                    for (int c = 0; c < G; ++c) {
                        const double w = It >= 0 ? sWG[c * R + K] : 0.0;
                        ar = fma(tur[c], w, ar);
                        ai = fma(tui[c], w, ai);
                    }
                    const int I0 = SH::state(nt, sl), I3 = SH::state(nt, 6 + sl);
                    const int cmin = I0 / REPS, cmax = (I3 >= 0 ? I3 : R - 1) / REPS;
                    const double ukr = px_pick<G>(tur, K / REPS, cmin, cmax), uki = px_pick<G>(tui, K / REPS, cmin, cmax);
                    const double sr = fma(mean, ukr, yr[m][nt][sl]);
                    const double si = fma(-mean, uki, yi[m][nt][sl]);
                    part += ar * sr - ai * si;
                }
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            const double mine = __shfl_sync(0xffffffffu, part, 4 * (lane & 7));   // row lane % 8 of this m-tile
            if ((lane >> 3) == mg + m) mytr = mine;
        }
    }
    if (live) bt.intensities[bt.job_out[job] + x] = finish_point(mytr, abs_scale, flags, ps, st, gx);
}

// ---- RawPatternPhase: a measured pattern resampled onto the specimen's 2-theta axis --------
// phases.py:97-104: interp1d(raw_x, raw_y, bounds_error=False, fill_value=0)(rad2deg(2 theta)),
// then LPF / machine correction like any other phase (phases.py:81-95, specimen.py:57-62).
// The pool entry of the phase holds raw_x (ascending, n values) followed by raw_y.
// scipy 1.18.1 _call_linear: hi = searchsorted(x, x_new) clipped to [1, n-1],
//   slope form  y = ((x_new - x_lo) / (x_hi - x_lo)) * y_hi + ((x_hi - x_new) / (x_hi - x_lo)) * y_lo
__global__ void __launch_bounds__(128) k_phase_raw(StaticDev st, BatchDev bt, const int *__restrict__ jobs) {
    __shared__ PhaseSmem ps_store;
    PhaseSmem *ps = &ps_store;
    const int job = jobs[blockIdx.x];
    const int pb = bt.job_pb[job], s = bt.job_spec[job];
    const int X = st.spec_npts[s];
    const int x = blockIdx.y * 128 + threadIdx.x;
    const int *pi = bt.phase_i + pb * 8;
    const int n = pi[2], flags = pi[4];
    if (threadIdx.x == 0) lpf_setup(ps, bt.phase_d[pb * 8], st.derived + s * 8);
    __syncthreads();
    if (x >= X) return;
    const double *rx = bt.matrices + pi[6], *ry = rx + n;
    const long long gx = st.spec_off[s] + x;
    const double xq = __dmul_rn(__dmul_rn(2.0, st.theta[gx]), 57.29577951308232);   // np.rad2deg(2 * theta)
    double v = 0.0;
    if (n >= 2 && xq >= rx[0] && xq <= rx[n - 1]) {
        int lo = 0, hi = n;                              // searchsorted, side = 'left': first index with rx[i] >= xq
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (rx[mid] < xq) lo = mid + 1; else hi = mid;
        }
        hi = min(max(lo, 1), n - 1);
        const double x_lo = rx[hi - 1], x_hi = rx[hi], den = __dadd_rn(x_hi, -x_lo);
        v = __dadd_rn(__dmul_rn(__ddiv_rn(__dadd_rn(xq, -x_lo), den), ry[hi]),
                      __dmul_rn(__ddiv_rn(__dadd_rn(x_hi, -xq), den), ry[hi - 1]));
    }
    bt.intensities[bt.job_out[job] + x] = finish_point(v, 1.0, flags, ps, st, gx);
}

// ---- any rank <= 64: vectors in shared memory (fallback, not tuned) --------------
__global__ void __launch_bounds__(PX_GEN_TPB) k_phase_generic(StaticDev st, BatchDev bt, const int *__restrict__ jobs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PhaseSmem *ps = reinterpret_cast<PhaseSmem *>(smem_raw);
    const int job = jobs[blockIdx.x];
    const int pb = bt.job_pb[job], s = bt.job_spec[job];
    const int X = st.spec_npts[s];
    const int x0 = blockIdx.y * PX_GEN_TPB;
    if (x0 >= X || phase_is_invalid(bt, pb, job, X, x0, PX_GEN_TPB)) return;
    const int *pi = bt.phase_i + pb * 8;
    const int G = pi[0], R = pi[1], REPS = R / G;
    double *sP = reinterpret_cast<double *>(smem_raw + sizeof(PhaseSmem));    // [R][R]
    double *sWG = sP + R * R;                                                // [G][R]
    double2 *vy = reinterpret_cast<double2 *>(sWG + G * R + ((R * R + G * R) & 1));   // [R][TPB] y, 16-byte aligned
    double2 *vw = vy + R * PX_GEN_TPB;                                       // [R][TPB] w
    double2 *vu = vw + R * PX_GEN_TPB;                                       // [G][TPB] u
    double2 *vp = vu + 8 * PX_GEN_TPB;                                       // [G][TPB] phi
    double *sCoef = reinterpret_cast<double *>(vp + 8 * PX_GEN_TPB);
    const double *px = bt.phase_x + pb * 4;
    const int ntop = (int)px[2];
    const double mean = px[0], abs_scale = px[1];
    const int flags = pi[4];

    load_phase_lists(bt, pb, G, ps, st.sumX, st.derived + s * 8);
    {
        const double *W = bt.matrices + pi[6];
        const double *P = W + R * R;
        for (int i = threadIdx.x; i < R * R; i += PX_GEN_TPB) sP[i] = P[i];
        for (int i = threadIdx.x; i < G * R; i += PX_GEN_TPB) {
            const int c = i / R, K = i % R;
            double acc = 0.0;
            for (int J = c * REPS; J < (c + 1) * REPS; ++J) acc += W[J * R + K];
            sWG[i] = acc;
        }
        const double *coef = bt.coef + (long long)pb * bt.ncap;
        for (int i = threadIdx.x; i <= ntop; i += PX_GEN_TPB) sCoef[i] = coef[i];
    }
    __syncthreads();
    const int x = x0 + threadIdx.x;
    if (x >= X) return;
    const int t = threadIdx.x;
    const long long gx = st.spec_off[s] + x;
    const double stl = st.stl[gx];
    for (int c = 0; c < G; ++c) {
        double a, b, p, q;
        component_factors(ps, c, stl, st.asf + gx, bt.layer_tab + gx, st.sumX, a, b, p, q);
        vu[c * PX_GEN_TPB + t] = make_double2(a, b);
        vp[c * PX_GEN_TPB + t] = make_double2(p, q);
    }
    for (int I = 0; I < R; ++I) vy[I * PX_GEN_TPB + t] = make_double2(0.0, 0.0);
    for (int n = ntop; n >= 1; --n) {
        const double cn = sCoef[n];
        for (int I = 0; I < R; ++I) {
            const int c = I / REPS;
            const double2 u = vu[c * PX_GEN_TPB + t], p = vp[c * PX_GEN_TPB + t], y = vy[I * PX_GEN_TPB + t];
            const double gr = p.x * u.x + p.y * u.y, gi = p.y * u.x - p.x * u.y;
            vw[I * PX_GEN_TPB + t] = make_double2(fma(p.x, y.x, fma(-p.y, y.y, cn * gr)),
                                                  fma(p.x, y.y, fma(p.y, y.x, cn * gi)));
        }
        for (int I = 0; I < R; ++I) {
            double ar = 0.0, ai = 0.0;
            for (int J = 0; J < R; ++J) {
                const double p = sP[I * R + J];
                const double2 w = vw[J * PX_GEN_TPB + t];
                ar = fma(p, w.x, ar);
                ai = fma(p, w.y, ai);
            }
            vy[I * PX_GEN_TPB + t] = make_double2(ar, ai);
        }
    }
    double tr = 0.0;
    for (int K = 0; K < R; ++K) {
        const int ck = K / REPS;
        double ar = 0.0, ai = 0.0;
        for (int c = 0; c < G; ++c) {
            const double w = sWG[c * R + K];
            const double2 u = vu[c * PX_GEN_TPB + t];
            ar = fma(u.x, w, ar);
            ai = fma(u.y, w, ai);
        }
        const double2 u = vu[ck * PX_GEN_TPB + t], y = vy[K * PX_GEN_TPB + t];
        tr += ar * fma(mean, u.x, y.x) - ai * fma(-mean, u.y, y.y);
    }
    bt.intensities[bt.job_out[job] + x] = finish_point(tr, abs_scale, flags, ps, st, gx);
}

// mmult (math_tools.py:16-17): C[x] = A[x] B[x] for X complex r x r matrices, in the order numpy evaluates the
// reference's broadcast-and-sum: every product on its own ((ar br - ai bi) + i (ar bi + ai br), no FMA -- numpy's SIMD
// loops fuse one of the two products on some CPUs, so the last bit is not portable there either), summed over the
// contraction index in ascending order starting from the first product.  get_Q_matrices (phases.py:41-46): Qn[0] = Q,
// Qn[n] = mmult(Qn[n-1], Q); a thread owns one ROW of one point's matrix: row i of Q^(n+1) only needs row i of Q^n.
__device__ __forceinline__ double2 px_cmul_np(double2 a, double2 b) {
    return make_double2(__dadd_rn(__dmul_rn(a.x, b.x), -__dmul_rn(a.y, b.y)), __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
}
__global__ void k_leaf_mmult(const double2 *__restrict__ A, const double2 *__restrict__ B, long long X, int r, double2 *__restrict__ Cout) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= X * r * r) return;
    const long long x = t / (r * r);
    const int i = (int)((t / r) % r), j = (int)(t % r);
    const double2 *a = A + (x * r + i) * r, *b = B + x * r * r + j;
    double2 acc = px_cmul_np(a[0], b[0]);
    for (int k = 1; k < r; ++k) {
        const double2 p = px_cmul_np(a[k], b[(long long)k * r]);
        acc = make_double2(__dadd_rn(acc.x, p.x), __dadd_rn(acc.y, p.y));
    }
    Cout[t] = acc;
}
__global__ void k_leaf_matrix_powers(const double2 *__restrict__ Q, long long X, int r, int nmax, double2 *__restrict__ Qn) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= X * r) return;
    const long long x = t / r;
    const int i = (int)(t % r);
    const double2 *q = Q + x * r * r;
    const long long level = X * r * r;                     // doubles2 per power
    double2 *row = Qn + (x * r + i) * r;                   // row i of Qn[0][x]
    for (int j = 0; j < r; ++j) row[j] = q[i * r + j];
    for (int n = 1; n <= nmax; ++n) {
        const double2 *prev = row + (long long)(n - 1) * level;
        double2 *next = row + (long long)n * level;
        for (int j = 0; j < r; ++j) {
            double2 acc = px_cmul_np(prev[0], q[j]);
            for (int k = 1; k < r; ++k) {
                const double2 p = px_cmul_np(prev[k], q[k * r + j]);
                acc = make_double2(__dadd_rn(acc.x, p.x), __dadd_rn(acc.y, p.y));
            }
            next[j] = acc;
        }
    }
}

// one component on an arbitrary stl axis (leaf: components.get_factors / atoms.get_structure_factor)
__global__ void k_leaf_component(const double *stl, long long n, const double *comp_d, int n_layer, int n_inter,
                                 const double *atom_d, const double *atype_rows, const unsigned char *has_type,
                                 double2 *sf, double2 *phi, double *z_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double lattice_d = comp_d[3];
    const double zf = __ddiv_rn(comp_d[0] - lattice_d, comp_d[1] - lattice_d);
    const double x = stl[i];
    const double v = __dmul_rn(x, 0.05);
    const double s = __dmul_rn(v, v);
    double sr = 0.0, si = 0.0;
    for (int a = 0; a < n_layer + n_inter; ++a) {
        const double z0 = atom_d[a * 2 + 1];
        const double z = (a < n_layer) ? z0 : __dadd_rn(lattice_d, __dmul_rn(zf, z0 - lattice_d));
        if (i == 0) z_out[a] = z;
        if (!has_type[a]) continue;
        plane_term(__dmul_rn(asf_value(atype_rows + a * 12, s), atom_d[a * 2]), __dmul_rn(PX_TWO_PI, z), x, sr, si);
    }
    sf[i] = make_double2(sr, si);
    double pr, pi_;
    phase_factor(x, comp_d[0], comp_d[2], pr, pi_);
    phi[i] = make_double2(pr, pi_);
}

// ---------------------------------------------------------------------------------
// wavelength distribution: I <- I + interp(asin(stl lambda_j / 2), f_j I)(theta), j in list order,
// each pass reading the pattern the previous pass produced (specimen.py:64-74).
// One launch per pass j, out of place (src -> dst, the caller ping-pongs two buffers): every
// point is independent, so the pass is a streaming gather with 4 points in flight per thread.
// The abscissa tables (index of the upper neighbour and the two interpolation weights) are
// candidate-invariant and L2-resident.
// ---------------------------------------------------------------------------------
#define PX_WL_TPB 256
#define PX_WL_PPT 4           // measured on c2: 4 points per thread 54 us, 2 (full occupancy) 58 us, 1: 82 us
__global__ void __launch_bounds__(PX_WL_TPB) k_wavelength_pass(StaticDev st, BatchDev bt, int job0, int pass,
                                                               const double *__restrict__ src, double *__restrict__ dst) {
    const int job = job0 + blockIdx.x;                  // jobs on x: grid.y stops at 65535
    const int s = bt.job_spec[job];
    const int X = st.spec_npts[s];
    const int xb = blockIdx.y * (PX_WL_TPB * PX_WL_PPT) + threadIdx.x;
    if (xb >= X) return;
    const long long out = bt.job_out[job];
    const double *I = src + out;
    double *O = dst + out;
    const bool active = bt.job_pb[job] >= 0 && pass < st.spec_nwl[s];     // a zero pattern stays zero
    const long long off = st.wl_off[s] + (long long)pass * X;
    const double f = active ? st.wl_fraction[st.wl_first[s] + pass] : 0.0;
    double v[PX_WL_PPT], hi_v[PX_WL_PPT], lo_v[PX_WL_PPT], whi[PX_WL_PPT], wlo[PX_WL_PPT];
    int hi[PX_WL_PPT];
#pragma unroll
    for (int u = 0; u < PX_WL_PPT; ++u) {
        const int x = xb + u * PX_WL_TPB;
        hi[u] = (active && x < X) ? __ldg(st.wl_index + off + x) : -1;
    }
#pragma unroll
    for (int u = 0; u < PX_WL_PPT; ++u) {
        const int x = xb + u * PX_WL_TPB;
        v[u] = (x < X) ? I[x] : 0.0;
        const bool g = hi[u] >= 0;
        hi_v[u] = g ? I[hi[u]] : 0.0;
        lo_v[u] = g ? I[hi[u] - 1] : 0.0;
        whi[u] = g ? __ldg(st.wl_w_hi + off + x) : 0.0;
        wlo[u] = g ? __ldg(st.wl_w_lo + off + x) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < PX_WL_PPT; ++u) {
        const int x = xb + u * PX_WL_TPB;
        if (x >= X) continue;
        double add = 0.0;
        if (hi[u] >= 0) {
            const double y_hi = __dmul_rn(hi_v[u], f), y_lo = __dmul_rn(lo_v[u], f);
            add = __dadd_rn(__dmul_rn(whi[u], y_hi), __dmul_rn(wlo[u], y_lo));
        }
        O[x] = __dadd_rn(v[u], add);
    }
}

// ---------------------------------------------------------------------------------
// mixture: total = scale * sum_p frac_p I_p + bg * C ; Rp = 100 sum|obs - total| / sum|obs|
// (specimen.py:84-88,16-19, statistics.py:22-27), optionally FUSED with the last pass of the wavelength distribution
// (GATHER): the block forms I_p <- I_p + interp(f_j I_p) for every phase of its tile on the fly -- the same operations
// in the same order as k_wavelength_pass -- writes the final pattern (dst) and feeds it straight into the sum, so the
// patterns are not read a second time and the abscissa tables (index + two weights, 20 B per point) are fetched once
// per point instead of once per point and phase.
// grid (candidates, tiles of PX_RES_TPB * PX_RES_PPT points, S): a block owns one tile of one (candidate, specimen) for all Z patterns and M
// phases, PX_RES_PPT points in flight per thread (one block per (candidate, specimen) looping over the whole pattern was
// latency-bound: 91 us against 50 us for c2).  Its sum |obs - total| goes to partial[((k - k0) S + s) ntiles + tile];
// k_residual_finish adds the tiles in order (deterministic).  want_totals = 0: the total pattern is only summed, not
// stored (callers that want residuals only).
// ---------------------------------------------------------------------------------
#define PX_RES_TPB 256
// two points per thread at full occupancy (8 blocks of 256 threads, <= 32 registers): the fused kernel is bound by the
// latency of its dependent loads (index -> gather -> sum), not by DRAM; measured on c2 (220 MB moved): 4 points / 80
// registers / 3 blocks 130 us, 4 / 64 / 4: 134, 2 / 6 blocks 129, 2 / 8 blocks 113, 1 / 8 blocks 185
#ifndef PX_RES_PPT
#define PX_RES_PPT 2
#define PX_RES_MINB 8
#endif
template <bool GATHER>
__global__ void __launch_bounds__(PX_RES_TPB, PX_RES_MINB) k_mixture_residual(StaticDev st, BatchDev bt, int k0, int want_scaled, int want_totals,
                                                                int pass, const double *__restrict__ src, double *__restrict__ dst,
                                                                double *__restrict__ partial) {
    const int s = blockIdx.z, k = k0 + blockIdx.x, tile = blockIdx.y;       // candidates on x: grid.y / z stop at 65535
    const int X = st.spec_npts[s], Z = st.Z, M = bt.M;
    const int xb = tile * (PX_RES_TPB * PX_RES_PPT) + threadIdx.x;
    double num = 0.0;
    if (tile * (PX_RES_TPB * PX_RES_PPT) < X) {
        const long long soff = st.spec_off[s];
        const long long cbase = (long long)k * bt.cand_stride + (long long)M * Z * soff;
        const double *I = src + cbase;
        double *O = GATHER ? dst + cbase : nullptr;
        double *scaled = (want_scaled && bt.scaled) ? bt.scaled + cbase : nullptr;
        double *tot = bt.totals + ((long long)k * st.sumX + soff) * Z;
        const double *obs = st.observed + (long long)Z * soff;
        const unsigned char *sel = st.selected + soff;
        const double *corr = st.corr + soff;
        const double scale = bt.scales[k * st.S + s], bg = bt.bgshifts[k * st.S + s];
        const double *frac = bt.fractions + (long long)k * M;
        const long long pstride = (long long)Z * X;                   // phase stride inside [M][Z][X]
        int hi[PX_RES_PPT];
        double whi[PX_RES_PPT], wlo[PX_RES_PPT], f = 0.0;
        if constexpr (GATHER) {
            const bool active = pass < st.spec_nwl[s];
            const long long off = st.wl_off[s] + (long long)pass * X;
            f = active ? st.wl_fraction[st.wl_first[s] + pass] : 0.0;
#pragma unroll
            for (int u = 0; u < PX_RES_PPT; ++u) {
                const int x = xb + u * PX_RES_TPB;
                hi[u] = (active && x < X) ? __ldg(st.wl_index + off + x) : -1;
                whi[u] = hi[u] >= 0 ? __ldg(st.wl_w_hi + off + x) : 0.0;
                wlo[u] = hi[u] >= 0 ? __ldg(st.wl_w_lo + off + x) : 0.0;
            }
        }
        for (int z = 0; z < Z; ++z) {
            const double *Iz = I + (long long)z * X;
            double *Sz = scaled ? scaled + (long long)z * X : nullptr;
            double acc[PX_RES_PPT];
#pragma unroll
            for (int u = 0; u < PX_RES_PPT; ++u) acc[u] = 0.0;
            for (int p = 0; p < M; ++p) {
                const double fp = frac[p];
                const double *Ip = Iz + p * pstride;
                double iv[PX_RES_PPT];
#pragma unroll
                for (int u = 0; u < PX_RES_PPT; ++u) iv[u] = (xb + u * PX_RES_TPB < X) ? Ip[xb + u * PX_RES_TPB] : 0.0;
                if constexpr (GATHER) {
                    double hv[PX_RES_PPT], lv[PX_RES_PPT];
#pragma unroll
                    for (int u = 0; u < PX_RES_PPT; ++u) {
                        hv[u] = hi[u] >= 0 ? Ip[hi[u]] : 0.0;
                        lv[u] = hi[u] >= 0 ? Ip[hi[u] - 1] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < PX_RES_PPT; ++u) {
                        if (hi[u] >= 0)             // the operations of k_wavelength_pass, in its order
                            iv[u] = __dadd_rn(iv[u], __dadd_rn(__dmul_rn(whi[u], __dmul_rn(hv[u], f)), __dmul_rn(wlo[u], __dmul_rn(lv[u], f))));
                        if (xb + u * PX_RES_TPB < X) O[(long long)z * X + p * pstride + xb + u * PX_RES_TPB] = iv[u];
                    }
                }
#pragma unroll
                for (int u = 0; u < PX_RES_PPT; ++u) {
                    const double v = __dmul_rn(__dmul_rn(fp, iv[u]), scale);
                    if (Sz && xb + u * PX_RES_TPB < X) Sz[p * pstride + xb + u * PX_RES_TPB] = v;
                    acc[u] = __dadd_rn(acc[u], v);
                }
            }
#pragma unroll
            for (int u = 0; u < PX_RES_PPT; ++u) {
                const int x = xb + u * PX_RES_TPB;
                if (x >= X) continue;
                const double t = __dadd_rn(acc[u], __dmul_rn(bg, corr[x]));
                if (want_totals) tot[(long long)z * X + x] = t;
                if (sel[x]) num += fabs(obs[(long long)z * X + x] - t);
            }
        }
    }
    __shared__ double red[32];
    for (int o = 16; o > 0; o >>= 1) num += __shfl_down_sync(0xffffffffu, num, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = num;
    __syncthreads();
    if (threadIdx.x < 32) {
        num = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) num += __shfl_down_sync(0xffffffffu, num, o);
        if (threadIdx.x == 0) partial[((long long)(k - k0) * st.S + s) * gridDim.y + tile] = num;
    }
}

// ---------------------------------------------------------------------------------
// ALL passes of the wavelength distribution + mixture + residual in one kernel (pyxrd_compute_all where the gathers of
// the distribution stay local).  A gather of pass j reaches at most r-_j points down and r+_j points up (Cu K-alpha2
// against K-alpha1 on the c2 grid: 21 points down), so the final value of a point depends on the raw pattern inside a
// window of HL = sum r-_j points below and HR = sum r+_j points above it (the host takes the reaches from the abscissa
// tables, pyxrd_set_static).  A block stages the raw pattern of its tile + halo in shared memory once per (z, phase),
// runs the passes between two shared buffers -- the operations of k_wavelength_pass / k_mixture_residual<true> in their
// order, bit for bit the same patterns -- and feeds the tile straight into the mixture sum.  The patterns cross HBM once
// (read, + halo) instead of four times (pass: read + write, residual: read + write); with PYXRD_RESIDUALS_ONLY nothing is
// written but the tile sums.  c2 (2048 candidates): wavelength pass + residual kernel 168 us -> 121 us, bound by instruction
// issue (measured: 4 blocks per SM at 64 registers beat 6 at 40 with spills; 8 candidates per block = 16 = 4).
// ---------------------------------------------------------------------------------
#define PX_FUSE_TPB 256
#define PX_FUSE_T 512                  // points of a tile
#define PX_FUSE_HMAX 128               // HL + HR the shared buffers hold; wider distributions keep the pass kernels
#define PX_FUSE_SLOTS ((PX_FUSE_T + PX_FUSE_HMAX + PX_FUSE_TPB - 1) / PX_FUSE_TPB)
#define PX_FUSE_MAXPASS 4
#ifndef PX_FUSE_MINB
#define PX_FUSE_MINB 4                 // blocks per SM the register allocation aims at
#endif
#ifndef PX_FUSE_CPB
#define PX_FUSE_CPB 8                  // candidates a block walks for its tile at most (grid.x = ceil(candidates / cpb); the host
#endif                                 // takes fewer where the population is too small to fill the machine with blocks)
// A block owns one tile of one specimen for PX_FUSE_CPB consecutive candidates: the abscissa tables of the tile (index of
// the upper neighbour and the two weights, per pass) are candidate-invariant and fetched once into registers; the raw
// pattern of the NEXT (candidate, z, phase) is prefetched into registers while the current one goes through the passes
// in shared memory, so that no global latency sits between the stages of a pattern.
template <int NPASS, bool STORE>         // STORE = false: residuals only, nothing but the tile sums is written
__global__ void __launch_bounds__(PX_FUSE_TPB, PX_FUSE_MINB) k_patterns_residual(StaticDev st, BatchDev bt, int k0, int k1, int cpb, int want_scaled, int want_totals,
                                                                  int HL, int HR, const double *__restrict__ src, double *__restrict__ dst,
                                                                  double *__restrict__ partial) {
    constexpr int NOUT = PX_FUSE_T / PX_FUSE_TPB, NBUF = PX_FUSE_SLOTS * PX_FUSE_TPB;    // every slot of every thread has a buffer entry
    const int s = blockIdx.z, tile = blockIdx.y, tid = threadIdx.x;
    const int kb = k0 + blockIdx.x * cpb, ke = min(k1, kb + cpb);
    const int X = st.spec_npts[s], Z = st.Z, M = bt.M, S = st.S;
    const int x0 = tile * PX_FUSE_T;
    __shared__ double bufR[NBUF], bufA[NPASS > 1 ? NBUF : 1], bufB[NPASS > 2 ? NBUF : 1];
    __shared__ double red[32];
    if (x0 >= X) {                                              // (specimens shorter than the longest one)
        if (tid == 0)
            for (int k = kb; k < ke; ++k) partial[((long long)(k - k0) * S + s) * gridDim.y + tile] = 0.0;
        return;
    }
    const int W = PX_FUSE_T + HL + HR, y0 = x0 - HL;
    const long long soff = st.spec_off[s];
    const long long pstride = (long long)Z * X;
    const int nwl = st.spec_nwl[s];
    const long long off0 = st.wl_off[s];
    const double *wf = st.wl_fraction + st.wl_first[s];
    // Tables of a pass at a buffer position: index of the upper neighbour INSIDE the buffer and the two weights.  No
    // neighbour (outside the pattern, fill value, a margin position nobody reads): index 1 with weights 0 -- the pass is
    // then v + (0 a + 0 b) = v without a branch (the same value as the pass kernels give; the sign of a zero may differ).
    auto table = [&](int j, int y, int &g, double &wh, double &wl) {
        g = 1; wh = 0.0; wl = 0.0;
        if (j < nwl && y >= 0 && y < X) {
            const long long off = off0 + (long long)j * X + y;
            const int hi = __ldg(st.wl_index + off);
            if (hi >= 0 && hi - 1 >= y0 && hi < y0 + W) { g = hi - y0; wh = __ldg(st.wl_w_hi + off); wl = __ldg(st.wl_w_lo + off); }
        }
    };
    int g0[PX_FUSE_SLOTS], gL[NOUT];
    double wh0[PX_FUSE_SLOTS], wl0[PX_FUSE_SLOTS], whL[NOUT], wlL[NOUT];
    unsigned in_mask = 0u;                                      // bit u: buffer position tid + u TPB lies inside the pattern
#pragma unroll
    for (int u = 0; u < PX_FUSE_SLOTS; ++u) {
        const int b = tid + u * PX_FUSE_TPB, y = y0 + b;
        table(0, b < W ? y : -1, g0[u], wh0[u], wl0[u]);
        if (b < W && y >= 0 && y < X) in_mask |= 1u << u;
    }
    // tile points of this thread: observed, background, mask (candidate-invariant when Z == 1; else read per z below)
    double corr_x[NOUT];
    bool live[NOUT], inside[NOUT];
#pragma unroll
    for (int u = 0; u < NOUT; ++u) {
        const int x = x0 + tid + u * PX_FUSE_TPB;
        inside[u] = x < X;
        table(NPASS - 1, inside[u] ? x : -1, gL[u], whL[u], wlL[u]);
        corr_x[u] = inside[u] ? st.corr[soff + x] : 0.0;
        live[u] = inside[u] && st.selected[soff + x] != 0;
    }
    const double *obs = st.observed + (long long)Z * soff + x0 + tid;
    const double f0 = 0 < nwl ? wf[0] : 0.0, fL = NPASS - 1 < nwl ? wf[NPASS - 1] : 0.0;
    // the patterns this block walks: candidates kb .. ke-1, per candidate z, then phase; the raw values of the NEXT pattern
    // are prefetched into registers while the current one goes through the passes (pointer arithmetic without divisions)
    const long long spec_base = (long long)M * Z * soff;
    double raw[PX_FUSE_SLOTS];
    auto prefetch = [&](const double *pat) {                    // pat: the pattern's point y0 + tid
#pragma unroll
        for (int u = 0; u < PX_FUSE_SLOTS; ++u) raw[u] = ((in_mask >> u) & 1u) ? pat[u * PX_FUSE_TPB] : 0.0;
    };
    const double *pat_next = src + (long long)kb * bt.cand_stride + spec_base + y0 + tid;
    prefetch(pat_next);
    for (int k = kb; k < ke; ++k) {
        const long long cbase = (long long)k * bt.cand_stride + spec_base + x0 + tid;
        double *O = (STORE && dst) ? dst + cbase : nullptr;
        double *scaled = (STORE && want_scaled && bt.scaled) ? bt.scaled + cbase : nullptr;
        double *tot = STORE ? bt.totals + ((long long)k * st.sumX + soff) * Z + x0 + tid : nullptr;
        const double scale = bt.scales[k * S + s], bg = bt.bgshifts[k * S + s];
        const double *frac = bt.fractions + (long long)k * M;
        double num = 0.0;
        for (int z = 0; z < Z; ++z) {
            double acc[NOUT];
#pragma unroll
            for (int u = 0; u < NOUT; ++u) acc[u] = 0.0;
            for (int p = 0; p < M; ++p) {
                const double fp = frac[p];
#pragma unroll
                for (int u = 0; u < PX_FUSE_SLOTS; ++u) bufR[tid + u * PX_FUSE_TPB] = raw[u];
                __syncthreads();
                // next pattern of the sequence: next phase, else the first phase of the next z, else the next candidate
                const bool more = p + 1 < M || z + 1 < Z || k + 1 < ke;
                if (p + 1 < M) pat_next += pstride;
                else if (z + 1 < Z) pat_next += X - (long long)(M - 1) * pstride;
                else pat_next += bt.cand_stride - (long long)(Z - 1) * X - (long long)(M - 1) * pstride;
                if (more) prefetch(pat_next);
                const double *cur = bufR;
                if constexpr (NPASS > 1) {
#pragma unroll
                    for (int u = 0; u < PX_FUSE_SLOTS; ++u) {   // first pass (operations of k_wavelength_pass, in its order)
                        const int b = tid + u * PX_FUSE_TPB;
                        const double add = __dadd_rn(__dmul_rn(wh0[u], __dmul_rn(bufR[g0[u]], f0)), __dmul_rn(wl0[u], __dmul_rn(bufR[g0[u] - 1], f0)));
                        bufA[b] = __dadd_rn(bufR[b], add);
                    }
                    __syncthreads();
                    cur = bufA;
                    double *nxt = bufB;
#pragma unroll
                    for (int j = 1; j + 1 < NPASS; ++j) {       // middle passes: tables on the fly
                        const double f = j < nwl ? wf[j] : 0.0;
                        double *out = nxt;
#pragma unroll
                        for (int u = 0; u < PX_FUSE_SLOTS; ++u) {
                            const int b = tid + u * PX_FUSE_TPB;
                            int g; double wh, wl;
                            table(j, b < W ? y0 + b : -1, g, wh, wl);
                            const double add = __dadd_rn(__dmul_rn(wh, __dmul_rn(cur[g], f)), __dmul_rn(wl, __dmul_rn(cur[g - 1], f)));
                            out[b] = __dadd_rn(cur[b], add);
                        }
                        __syncthreads();
                        nxt = const_cast<double *>(cur); cur = out;
                    }
                }
                double iv[NOUT];
#pragma unroll
                for (int u = 0; u < NOUT; ++u) {                // last pass, tile points only
                    const double add = __dadd_rn(__dmul_rn(whL[u], __dmul_rn(cur[gL[u]], fL)), __dmul_rn(wlL[u], __dmul_rn(cur[gL[u] - 1], fL)));
                    iv[u] = __dadd_rn(cur[HL + tid + u * PX_FUSE_TPB], add);
                }
                const long long poff = (long long)z * X + p * pstride;
#pragma unroll
                for (int u = 0; u < NOUT; ++u) {
                    if (STORE && O && inside[u]) O[poff + u * PX_FUSE_TPB] = iv[u];
                    const double v = __dmul_rn(__dmul_rn(fp, iv[u]), scale);
                    if (STORE && scaled && inside[u]) scaled[poff + u * PX_FUSE_TPB] = v;
                    acc[u] = __dadd_rn(acc[u], v);
                }
                if (NPASS == 1) __syncthreads();                // (bufR is read by the single pass and rewritten by the next pattern)
            }
#pragma unroll
            for (int u = 0; u < NOUT; ++u) {
                const double t = __dadd_rn(acc[u], __dmul_rn(bg, corr_x[u]));
                if (STORE && want_totals && inside[u]) tot[(long long)z * X + u * PX_FUSE_TPB] = t;
                const double o = inside[u] ? obs[(long long)z * X + u * PX_FUSE_TPB] : 0.0;
                if (live[u]) num += fabs(o - t);
            }
        }
        // tile sum of this candidate (the reduction of k_mixture_residual)
        for (int o = 16; o > 0; o >>= 1) num += __shfl_down_sync(0xffffffffu, num, o);
        if ((tid & 31) == 0) red[tid >> 5] = num;
        __syncthreads();
        if (tid < 32) {
            num = tid < (PX_FUSE_TPB >> 5) ? red[tid] : 0.0;
            for (int o = 16; o > 0; o >>= 1) num += __shfl_down_sync(0xffffffffu, num, o);
            if (tid == 0) partial[((long long)(k - k0) * S + s) * gridDim.y + tile] = num;
        }
        __syncthreads();                                        // red[] is rewritten by the next candidate
    }
}

// residual of every (candidate, specimen) from the tile sums, in tile order
__global__ void k_residual_finish(StaticDev st, BatchDev bt, int k0, int k1, int ntiles, const double *__restrict__ partial) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int S = st.S;
    if (i >= (k1 - k0) * S) return;
    const int k = k0 + i / S, s = i % S;
    double num = 0.0;
    for (int t = 0; t < ntiles; ++t) num += partial[(long long)i * ntiles + t];
    const double den = st.derived[s * 8 + 3];
    const int n = st.Z * st.spec_npts[s];
    const bool has_obs = st.gonio[s * 12 + 11] != 0.0;                                    // no observations -> 0.0 (mixture.py:247-251)
    bt.residuals[(long long)k * (S + 1) + 1 + s] = (n > 0 && has_obs) ? num / den * 100.0 : 0.0;
}

// k_residual_finish + k_mixture_mean in one launch (Rp: nothing runs between them): thread per candidate, the tile sums of
// its specimens in tile order, then the mean over the specimens -- the operations of the two kernels in their order
__global__ void k_residual_finish_mean(StaticDev st, BatchDev bt, int k0, int k1, int ntiles, const double *__restrict__ partial) {
    const int k = k0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k1) return;
    const int S = st.S;
    double acc = 0.0;
    for (int s = 0; s < S; ++s) {
        const double *pk = partial + ((long long)(k - k0) * S + s) * ntiles;
        double num = 0.0;
        for (int t = 0; t < ntiles; ++t) num += pk[t];
        const double den = st.derived[s * 8 + 3];
        const int n = st.Z * st.spec_npts[s];
        const bool has_obs = st.gonio[s * 12 + 11] != 0.0;                                    // no observations -> 0.0 (mixture.py:247-251)
        const double r = (n > 0 && has_obs) ? num / den * 100.0 : 0.0;
        bt.residuals[(long long)k * (S + 1) + 1 + s] = r;
        acc += r;
    }
    bt.residuals[(long long)k * (S + 1)] = acc / S;                                           // np.average(residuals[1:])
}

// ---------------------------------------------------------------------------------
// Rpder (settings.RESIDUAL_METHOD = "Rpder", statistics.py:29-48, math_tools.py:46-99):
//   Rp(gradient(smooth(exp, 15)), gradient(smooth(calc, 15))) on the CLIPPED patterns, i.e. the
//   selected points of all Z patterns flattened into one vector (smooth() flattens, :84-85).
//   smooth = 31-tap normalised Blackman window over s = [x[30:0:-1], x, x[-1:-31:-1]] ('valid'
//   convolution, then [15:-15]):  y[i] = sum_k w[k] xr(i + 15 - k),
//   xr(t) = x[-t] (t < 0), x[t], x[2n - 1 - t] (t >= n)  -- the right reflection repeats the edge.
//   np.gradient: (y[i+1] - y[i-1]) / 2 inside, one-sided differences at the two ends.
// One block per (specimen, candidate); the clipped exp / calc vectors are staged in shared
// memory when they fit (use_smem), else gathered from global memory.
// ---------------------------------------------------------------------------------
__constant__ double PX_BLACKMAN[31];      // np.blackman(31) / sum, written once by pyxrd_create

__global__ void k_selected_positions(StaticDev st, int *sel_pos, int *sel_count) {
    const int s = blockIdx.x;
    if (threadIdx.x != 0) return;
    const int X = st.spec_npts[s];
    const unsigned char *sel = st.selected + st.spec_off[s];
    int *out = sel_pos + st.spec_off[s];
    int n = 0;
    for (int x = 0; x < X; ++x)
        if (sel[x]) out[n++] = x;
    sel_count[s] = n;
}

template <bool SMEM>
__device__ __forceinline__ double rpder_value(const double *v, const double *pat, const int *pos, int nsel, int X, int j) {
    if (SMEM) return v[j];
    const int z = j / nsel;
    return pat[(long long)z * X + pos[j - z * nsel]];
}

template <bool SMEM>
__device__ __forceinline__ double rpder_smooth(const double *v, const double *pat, const int *pos, int nsel, int X, int n, int i) {
    double acc = 0.0;
#pragma unroll 1
    for (int k = 0; k < 31; ++k) {
        const int t = i + 15 - k;
        const int j = t < 0 ? -t : (t >= n ? 2 * n - 1 - t : t);
        acc = fma(PX_BLACKMAN[k], rpder_value<SMEM>(v, pat, pos, nsel, X, j), acc);
    }
    return acc;
}

template <bool SMEM>
__device__ __forceinline__ double rpder_gradient(const double *v, const double *pat, const int *pos, int nsel, int X, int n, int i) {
    const int a = i > 0 ? i - 1 : 0, b = i < n - 1 ? i + 1 : n - 1;
    const double d = rpder_smooth<SMEM>(v, pat, pos, nsel, X, n, b) - rpder_smooth<SMEM>(v, pat, pos, nsel, X, n, a);
    return (i > 0 && i < n - 1) ? d / 2.0 : d;
}

template <bool SMEM>
__global__ void __launch_bounds__(256) k_residual_rpder(StaticDev st, BatchDev bt, int k0, const int *sel_pos, const int *sel_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int s = blockIdx.y, k = k0 + blockIdx.x;      // candidates on x: grid.y stops at 65535
    const int X = st.spec_npts[s], Z = st.Z;
    const int nsel = sel_count[s], n = Z * nsel;
    const long long soff = st.spec_off[s];
    const double *tot = bt.totals + ((long long)k * st.sumX + soff) * Z;
    const double *obs = st.observed + (long long)Z * soff;
    const int *pos = sel_pos + soff;
    double *ve = reinterpret_cast<double *>(smem_raw), *vc = ve + n;
    if (SMEM) {
        for (int j = threadIdx.x; j < n; j += blockDim.x) {
            const int z = j / nsel, x = pos[j - z * nsel];
            ve[j] = obs[(long long)z * X + x];
            vc[j] = tot[(long long)z * X + x];
        }
        __syncthreads();
    }
    double num = 0.0, den = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double de = rpder_gradient<SMEM>(ve, obs, pos, nsel, X, n, i);
        const double dc = rpder_gradient<SMEM>(vc, tot, pos, nsel, X, n, i);
        num += fabs(de - dc);
        den += fabs(de);
    }
    __shared__ double red[2][32];
    for (int o = 16; o > 0; o >>= 1) {
        num += __shfl_down_sync(0xffffffffu, num, o);
        den += __shfl_down_sync(0xffffffffu, den, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = num; red[1][threadIdx.x >> 5] = den; }
    __syncthreads();
    if (threadIdx.x < 32) {
        num = threadIdx.x < (blockDim.x >> 5) ? red[0][threadIdx.x] : 0.0;
        den = threadIdx.x < (blockDim.x >> 5) ? red[1][threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) {
            num += __shfl_down_sync(0xffffffffu, num, o);
            den += __shfl_down_sync(0xffffffffu, den, o);
        }
        if (threadIdx.x == 0)
            bt.residuals[(long long)k * (st.S + 1) + 1 + s] = (Z * X > 0 && st.gonio[s * 12 + 11] != 0.0) ? num / den * 100.0 : 0.0;
    }
}

__global__ void k_mixture_mean(int k0, int K, int S, double *residuals) {
    const int k = k0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    double *r = residuals + (long long)k * (S + 1);
    double acc = 0.0;
    for (int s = 0; s < S; ++s) acc += r[1 + s];
    r[0] = acc / S;                                           // np.average(residuals[1:])
}

// ---------------------------------------------------------------------------------
// residual rows of this rank into the gathered table of EVERY rank of the box (the only exchange of the
// path).  The peers' tables are mapped into this process over NVLink / NVSwitch (torch symmetric memory or CUDA IPC, the
// caller's business); the kernel is plain stores to peer memory -- 8 B per element and peer, latency-bound -- followed by
// a system-scope fence; the caller's cross-GPU barrier orders them against the peers' reads.
// ---------------------------------------------------------------------------------
#define PX_MAX_PEERS 16
struct PeerTables { double *table[PX_MAX_PEERS]; int n; };
__global__ void k_put_rows_to_peers(const double *__restrict__ src, long long count, long long dst0, PeerTables peers) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) {
        const double v = src[i];
#pragma unroll 1
        for (int p = 0; p < peers.n; ++p) peers.table[p][dst0 + i] = v;
    }
    __threadfence_system();
}

// ---------------------------------------------------------------------------------
// FP64 FMA peak microbenchmark: 8 independent chains per thread
// ---------------------------------------------------------------------------------
__global__ void k_dfma_peak(double *out, int iters, double seed) {
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// ---------------------------------------------------------------------------------
// pyxrd/calculations/statistics.py on plain arrays (the GUI's pattern statistics, specimen/models/statistics.py:16,
// 115-118): one block per call.
//   kind 0  Rp          100 sum|e - c| / sum|e|                                   statistics.py:22-27
//   kind 1  Rpw         100 sqrt( sum' (e - c)^2 / e  /  sum' |e| ), terms that are nan / inf skipped; 0 when the
//                       quotient cannot be formed                                  statistics.py:50-68
//   kind 2  R_squared   1 - sum (e - c)^2 / sum (e - mean e)^2                    statistics.py:12-20
//   kind 3  Rpe         100 sqrt( (n - extra) / sum e^2 )                          statistics.py:70-79
//   kind 4  Rphase      100 sqrt( sum' (e - c)^2 p / e^2  /  sum' |p| )            statistics.py:81-100
// ---------------------------------------------------------------------------------
__device__ __forceinline__ double leaf_block_sum(double v, double *red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];      // same order in every thread
    return t;
}

__global__ void __launch_bounds__(256) k_leaf_statistic(int kind, const double *__restrict__ e, const double *__restrict__ c,
                                                        const double *__restrict__ ph, long long n, double extra, double *out) {
    __shared__ double red[8];
    double a = 0.0, b = 0.0;
    if (kind == 2) {                                     // mean first
        double sum = 0.0;
        for (long long i = threadIdx.x; i < n; i += blockDim.x) sum += e[i];
        const double avg = leaf_block_sum(sum, red) / (double)n;
        for (long long i = threadIdx.x; i < n; i += blockDim.x) {
            const double d = e[i] - c[i], m = e[i] - avg;
            a = fma(d, d, a);
            b = fma(m, m, b);
        }
    } else {
        for (long long i = threadIdx.x; i < n; i += blockDim.x) {
            const double ei = e[i], d = ei - (c ? c[i] : 0.0);
            if (kind == 0) { a += fabs(d); b += fabs(ei); }
            else if (kind == 1) {
                const double t = __ddiv_rn(__dmul_rn(d, d), ei);
                if (!(isnan(t) || isinf(t))) { a += t; b += fabs(ei); }
            } else if (kind == 3) { b = fma(ei, ei, b); }
            else {
                const double t = __ddiv_rn(__dmul_rn(__dmul_rn(d, d), ph[i]), __dmul_rn(ei, ei));
                if (!(isnan(t) || isinf(t))) { a += t; b += fabs(ph[i]); }
            }
        }
    }
    a = leaf_block_sum(a, red);
    b = leaf_block_sum(b, red);
    if (threadIdx.x == 0) {
        double r;
        if (kind == 0) r = a / b * 100.0;
        else if (kind == 1 || kind == 4) { const double q = a / b; r = (b != 0.0 && q >= 0.0) ? sqrt(q) * 100.0 : 0.0; }
        else if (kind == 2) r = 1.0 - a / b;
        else r = sqrt(((double)n - extra) / b) * 100.0;
        out[0] = r;
    }
}

// derive() = np.gradient(smooth(pattern, 15)) (statistics.py:29-41, math_tools.py:46-99); smooth_only: smooth_pattern()
__global__ void k_leaf_derive(const double *__restrict__ v, int n, int smooth_only, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = smooth_only ? rpder_smooth<true>(v, nullptr, nullptr, 0, 0, n, i) : rpder_gradient<true>(v, nullptr, nullptr, 0, 0, n, i);
}

// ---------------------------------------------------------------------------------
// K4: device-resident mixture optimiser, one CTA per candidate.
//
// Minimises the reference's objective (mixture.py:166-184: mean over specimens of
// Rp_s = 100 sum|obs - scale_s sum_p frac_p I_sp - bg_s C_s| / sum|obs|, statistics.py:22-27)
// over fractions >= 0, scales > 0, bgshifts >= 0 by majorise-minimise (iteratively reweighted
// least squares): with weights 1/max(|r_i|, delta) the L1 objective is majorised by a weighted
// least-squares problem whose data enter only through per-specimen Gram matrices
// B_s = sum_i w_i phi_i phi_i^T, c_s = sum_i w_i phi_i obs_i, phi_i = (I_s1..I_sM, C_s)(i).
// One pass over the candidate's patterns builds B_s, c_s (thread per point, block reduction);
// the small bilinear problem in v = (f[M], rho[S], bg[S]), theta_s = (rho_s f, bg_s), is then
// solved by a projected Levenberg-Marquardt / Newton iteration on thread 0.
// Parametrisation: unnormalised fractions f with per-specimen factors rho_s; the reference's
// (fractions, scales) are f / sum f and rho_s * sum f (mixture.py:32-48,196-205).
// ---------------------------------------------------------------------------------
#define PX_OPT_MAXS 8
#define PX_OPT_MAXMB 12
#define PX_OPT_MAXD (PX_OPT_MAXMB - 1 + 2 * PX_OPT_MAXS)
#define PX_OPT_TPB 256

struct OptParams {
    int outer;          // majorise-minimise iterations
    int lm_iters;       // Newton / LM iterations per majoriser
    int refine_bg;      // 0: bgshifts stay at the values of the batch
    int single_from;    // from this majoriser on ONE Newton / LM iteration per majoriser (0: lm_iters throughout)
    double delta0, shrink, delta_min;   // weight floor relative to mean |obs|: delta0 * shrink^it >= delta_min
    double newton_tol;                  // relative decrease of the majoriser below which no further Newton iteration is taken
    int coarse_until, coarse_stride;    // passes before coarse_until visit every coarse_stride-th group of points
    int lane_solver;                    // 5 < D <= 12: dampings of the single-lane Levenberg-Marquardt solver (0: warp teams)
};

#ifndef PX_OPT_NDAMP
#define PX_OPT_NDAMP 4     // warps (= dampings tried side by side) of the register Levenberg-Marquardt solver
#endif
#ifndef PX_OPT_PTR_MB
#define PX_OPT_PTR_MB 4      // up to this many columns the Gram pass reads through running pointers
#endif
#define PX_OPT_LDA 29      // row stride (doubles) of the d x d work matrices: odd, so a column is spread over all banks

struct OptSmem {
    double v[PX_OPT_MAXD];
    double lo[PX_OPT_MAXD];
    double B[PX_OPT_MAXS][PX_OPT_MAXMB * PX_OPT_MAXMB];
    double c[PX_OPT_MAXS][PX_OPT_MAXMB];
    double wgt[PX_OPT_MAXS];        // 100 / (S * sum|obs|)
    double mean_obs[PX_OPT_MAXS];
    double obj[PX_OPT_MAXS];        // sum |r| at the solution the last pass was taken at
    double H[PX_OPT_MAXD * PX_OPT_LDA];
    double A[PX_OPT_MAXD * PX_OPT_LDA];
    double g[PX_OPT_MAXD], vn[PX_OPT_MAXD], invd[PX_OPT_MAXD];
    double lam;                     // Levenberg-Marquardt damping, carried from one majoriser to the next
    double newton_tol;              // a Newton / LM iteration that lowers the majoriser by less than this (relative) is the last
    int lane_solver;                // dampings of the single-lane solver (0: the warp-team solvers)
    double cand_q[PX_OPT_NDAMP];                    // majoriser value each warp's damping reaches (register solver)
    double cand_v[PX_OPT_NDAMP][PX_OPT_MAXD];       // and its trial point
#ifdef PX_OPT_PROFILE
    long long prof[8];              // cycles: gram, reduce, newton, assembly, chol, solve+q; counts: tries, lm iterations
#endif
    int idx[PX_OPT_MAXD];
    double red[PX_OPT_TPB / 32][PX_OPT_MAXMB * (PX_OPT_MAXMB + 1) / 2 + PX_OPT_MAXMB + 1];
};

__device__ __forceinline__ double opt_warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}

#ifdef PX_OPT_PROFILE
__device__ __forceinline__ unsigned __smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
#endif

// The small bilinear problem is solved by warp 0 as a team: lane p owns fraction p / row p of the
// work matrices.  Every function below is called by all 32 lanes and returns warp-uniform values.

// per-specimen quadratic forms f^T B f, B[:,M]^T f, c^T f at v; lane p < M also gets t = (B f)[p]
__device__ __forceinline__ void opt_forms(const OptSmem &sm, const double *v, int s, int M, int MB, int lane,
                                          double &t, double &fBf, double &Bcf, double &cf) {
    const double *B = sm.B[s], *c = sm.c[s];
    t = 0.0;
    double a = 0.0, b = 0.0, d = 0.0;
    if (lane < M) {
        for (int q = 0; q < M; ++q) t = fma(B[lane * MB + q], v[q], t);
        a = v[lane] * t;
        b = B[lane * MB + M] * v[lane];
        d = c[lane] * v[lane];
    }
    // the three butterflies interleaved: 5 dependent steps instead of 15
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double a2 = __shfl_xor_sync(0xffffffffu, a, o), b2 = __shfl_xor_sync(0xffffffffu, b, o),
                     d2 = __shfl_xor_sync(0xffffffffu, d, o);
        a += a2; b += b2; d += d2;
    }
    fBf = a;
    Bcf = b;
    cf = d;
}

// majoriser value (up to a constant) at v
__device__ double opt_q(const OptSmem &sm, const double *v, int M, int S, int MB, int lane) {
    double q = 0.0;
    for (int s = 0; s < S; ++s) {
        double t, fBf, Bcf, cf;
        opt_forms(sm, v, s, M, MB, lane, t, fBf, Bcf, cf);
        const double r = v[M + s], b = v[M + S + s];
        q += sm.wgt[s] * (r * r * fBf + 2.0 * r * b * Bcf + sm.B[s][M * MB + M] * b * b - 2.0 * r * cf - 2.0 * sm.c[s][M] * b);
    }
    return q;
}

// Cholesky factor of the n x n matrix A (lower triangle, in place; lane i owns row i), reciprocal
// diagonal in invd; false if A is not positive definite
__device__ bool opt_chol(double *A, double *invd, int n, int lane) {
    for (int j = 0; j < n; ++j) {
        double a = 0.0;
        if (lane >= j && lane < n) {
            a = A[lane * PX_OPT_LDA + j];
            for (int k = 0; k < j; ++k) a = fma(-A[lane * PX_OPT_LDA + k], A[j * PX_OPT_LDA + k], a);
        }
        const double dj = __shfl_sync(0xffffffffu, a, j);
        if (!(dj > 0.0)) return false;
        const double inv = rsqrt(dj), sd = dj * inv;       // one reciprocal square root instead of sqrt + division
        if (lane == j) { A[j * PX_OPT_LDA + j] = sd; invd[j] = inv; }
        else if (lane > j && lane < n) A[lane * PX_OPT_LDA + j] = a * inv;
        __syncwarp();
    }
    return true;
}

// solves L L^T x = y; lane i passes y_i and receives x_i
__device__ double opt_chol_solve(const double *A, const double *invd, int n, int lane, double y) {
    for (int j = 0; j < n; ++j) {
        const double zj = __shfl_sync(0xffffffffu, y, j) * invd[j];
        if (lane == j) y = zj;
        else if (lane > j && lane < n) y = fma(-A[lane * PX_OPT_LDA + j], zj, y);
    }
    for (int j = n - 1; j >= 0; --j) {
        const double xj = __shfl_sync(0xffffffffu, y, j) * invd[j];
        if (lane == j) y = xj;
        else if (lane < j) y = fma(-A[j * PX_OPT_LDA + lane], xj, y);
    }
    return y;
}

// projected Levenberg-Marquardt / Newton on the majoriser (warp 0)
__device__ __noinline__ void opt_inner_solve(OptSmem &sm, int M, int S, int MB, int lm_iters, int refine_bg, int lane) {
    const int d = M + 2 * S;
    double lam = sm.lam;
    double *v = sm.v, *g = sm.g, *H = sm.H;
    for (int it = 0; it < lm_iters; ++it) {
#ifdef PX_OPT_PROFILE
        long long tA = clock64();
#endif
        for (int i = lane; i < d * PX_OPT_LDA; i += 32) H[i] = 0.0;
        __syncwarp();
        double gp = 0.0, qv = 0.0;                      // qv: majoriser value at v, a by-product of the forms below
        for (int s = 0; s < S; ++s) {
            const double *B = sm.B[s], *c = sm.c[s];
            const double w = sm.wgt[s], r = v[M + s], b = v[M + S + s];
            double t, fBf, Bcf, cf;
            opt_forms(sm, v, s, M, MB, lane, t, fBf, Bcf, cf);
            qv += w * (r * r * fBf + 2.0 * r * b * Bcf + B[M * MB + M] * b * b - 2.0 * r * cf - 2.0 * c[M] * b);
            const int ir = M + s, ib = M + S + s;
            if (lane < M) {
                const int p = lane;
                gp += w * (2.0 * r * r * t + 2.0 * r * b * B[p * MB + M] - 2.0 * r * c[p]);
                for (int q = 0; q < M; ++q) H[p * PX_OPT_LDA + q] += 2.0 * w * r * r * B[p * MB + q];
                const double hfr = w * (4.0 * r * t + 2.0 * b * B[p * MB + M] - 2.0 * c[p]);
                H[p * PX_OPT_LDA + ir] = hfr;
                H[ir * PX_OPT_LDA + p] = hfr;
                const double hfb = 2.0 * w * r * B[p * MB + M];
                H[p * PX_OPT_LDA + ib] = hfb;
                H[ib * PX_OPT_LDA + p] = hfb;
            }
            if (lane == 0) {
                g[ir] = w * (2.0 * r * fBf + 2.0 * b * Bcf - 2.0 * cf);
                g[ib] = w * (2.0 * r * Bcf + 2.0 * B[M * MB + M] * b - 2.0 * c[M]);
                H[ir * PX_OPT_LDA + ir] = 2.0 * w * fBf;
                H[ir * PX_OPT_LDA + ib] = H[ib * PX_OPT_LDA + ir] = 2.0 * w * Bcf;
                H[ib * PX_OPT_LDA + ib] = 2.0 * w * B[M * MB + M];
            }
        }
        if (lane < M) g[lane] = gp;
        __syncwarp();
        // free variables: not a fixed background, not held at its bound by the gradient
        bool free_i = false;
        if (lane < d) {
            const bool fixed_bg = (!refine_bg && lane >= M + S);
            const bool at_bound = (v[lane] <= sm.lo[lane] && g[lane] > 0.0);
            free_i = !fixed_bg && !at_bound && H[lane * PX_OPT_LDA + lane] > 0.0;
        }
        const unsigned mask = __ballot_sync(0xffffffffu, free_i);
        const int n = __popc(mask);
        if (free_i) sm.idx[__popc(mask & ((1u << lane) - 1u))] = lane;
        __syncwarp();
        if (n == 0) break;
#ifdef PX_OPT_PROFILE
        if (lane == 0) { sm.prof[3] += clock64() - tA; sm.prof[7] += 1; }
#endif
        bool accepted = false;
        double rel = 0.0;
        for (int tr = 0; tr < 8 && !accepted; ++tr) {
            if (lane < n) {
                const int ia = sm.idx[lane];
                for (int bcol = 0; bcol < n; ++bcol) sm.A[lane * PX_OPT_LDA + bcol] = H[ia * PX_OPT_LDA + sm.idx[bcol]];
                sm.A[lane * PX_OPT_LDA + lane] += lam * fabs(sm.A[lane * PX_OPT_LDA + lane]) + 1e-300;
            }
            __syncwarp();
#ifdef PX_OPT_PROFILE
            long long tC = clock64();
            if (lane == 0) sm.prof[6] += 1;
#endif
            if (!opt_chol(sm.A, sm.invd, n, lane)) { lam *= 10.0; __syncwarp(); continue; }
#ifdef PX_OPT_PROFILE
            if (lane == 0) sm.prof[4] += clock64() - tC;
            tC = clock64();
#endif
            const double dx = opt_chol_solve(sm.A, sm.invd, n, lane, lane < n ? -g[sm.idx[lane]] : 0.0);
            if (lane < d) sm.vn[lane] = v[lane];
            __syncwarp();
            if (lane < n) sm.vn[sm.idx[lane]] = fmax(sm.lo[sm.idx[lane]], v[sm.idx[lane]] + dx);
            __syncwarp();
            const double qn = opt_q(sm, sm.vn, M, S, MB, lane);
            if (qn < qv) {
                accepted = true;
                rel = (qv - qn) / fmax(fabs(qv), 1e-300);
                if (lane < d) v[lane] = sm.vn[lane];
                lam = fmax(lam * 0.2, 1e-12);
            } else {
                lam *= 5.0;
            }
            __syncwarp();
#ifdef PX_OPT_PROFILE
            if (lane == 0) sm.prof[5] += clock64() - tC;
#endif
        }
        if (!accepted || rel < sm.newton_tol) break;
    }
    if (lane == 0) sm.lam = fmin(lam, 1e-3);            // never start a majoriser more damped than the first one
}

// The same projected Levenberg-Marquardt step with everything in registers (M, S compile-time, D = M + 2 S <= 12):
// lane i owns variable i, its gradient entry and row i of the (symmetric) matrices.  Variables that are not free
// (fixed background, held at a bound by the gradient) get an identity row / column and a zero gradient instead of
// being compacted away, so every index is static.  Right-looking Cholesky: column j is scaled by rsqrt(d_jj)
// and broadcast entry by entry (independent shuffles, no dependent LDS chain); lane j keeps the broadcast column
// as row j of L^T for the backward substitution.
// barrier of the PX_OPT_NDAMP solver warps (named barrier 1; the rest of the block waits at barrier 0)
__device__ __forceinline__ void opt_team_sync() { asm volatile("bar.sync 1, %0;" ::"n"(PX_OPT_NDAMP * 32) : "memory"); }

template <int M, int S>
__device__ __forceinline__ double opt_q_reg(const OptSmem &sm, double vi, int lane) {
    constexpr int MB = M + 1;
    double q = 0.0;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const double *B = sm.B[s], *c = sm.c[s];
        const double r = __shfl_sync(0xffffffffu, vi, M + s), b = __shfl_sync(0xffffffffu, vi, M + S + s);
        double t = 0.0;
#pragma unroll
        for (int q2 = 0; q2 < M; ++q2) {
            const double fq = __shfl_sync(0xffffffffu, vi, q2);
            if (lane < M) t = fma(B[lane * MB + q2], fq, t);
        }
        double a = 0.0, bb = 0.0, dd = 0.0;
        if (lane < M) { a = vi * t; bb = B[lane * MB + M] * vi; dd = c[lane] * vi; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double a2 = __shfl_xor_sync(0xffffffffu, a, o), b2 = __shfl_xor_sync(0xffffffffu, bb, o),
                         d2 = __shfl_xor_sync(0xffffffffu, dd, o);
            a += a2; bb += b2; dd += d2;
        }
        q += sm.wgt[s] * (r * r * a + 2.0 * r * b * bb + B[M * MB + M] * b * b - 2.0 * r * dd - 2.0 * c[M] * b);
    }
    return q;
}

template <int M, int S>
__device__ __noinline__ void opt_inner_solve_reg(OptSmem &sm, int lm_iters, int refine_bg, int lane, int warp) {
    constexpr int MB = M + 1, D = M + 2 * S, NW = PX_OPT_NDAMP;
    double lam = sm.lam;
    double vi = lane < D ? sm.v[lane] : 0.0;
    const double lo_i = lane < D ? sm.lo[lane] : 0.0;
    const bool fixed_bg = (!refine_bg && lane >= M + S);
    for (int it = 0; it < lm_iters; ++it) {
#ifdef PX_OPT_PROFILE
        long long tA = clock64();
#endif
        double h[D];
#pragma unroll
        for (int j = 0; j < D; ++j) h[j] = 0.0;
        double gi = 0.0, qv = 0.0;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const double *B = sm.B[s], *c = sm.c[s];
            const double w = sm.wgt[s];
            const double r = __shfl_sync(0xffffffffu, vi, M + s), b = __shfl_sync(0xffffffffu, vi, M + S + s);
            double brow[M], t = 0.0, bm = 0.0, cl = 0.0;
#pragma unroll
            for (int q = 0; q < M; ++q) {
                const double fq = __shfl_sync(0xffffffffu, vi, q);
                brow[q] = lane < M ? B[lane * MB + q] : 0.0;
                t = fma(brow[q], fq, t);
            }
            if (lane < M) { bm = B[lane * MB + M]; cl = c[lane]; }
            double a = lane < M ? vi * t : 0.0, bb = lane < M ? bm * vi : 0.0, dd = lane < M ? cl * vi : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double a2 = __shfl_xor_sync(0xffffffffu, a, o), b2 = __shfl_xor_sync(0xffffffffu, bb, o),
                             d2 = __shfl_xor_sync(0xffffffffu, dd, o);
                a += a2; bb += b2; dd += d2;
            }
            const double fBf = a, Bcf = bb, cf = dd, BMM = B[M * MB + M], cM = c[M];
            qv += w * (r * r * fBf + 2.0 * r * b * Bcf + BMM * b * b - 2.0 * r * cf - 2.0 * cM * b);
            const double hfr = w * (4.0 * r * t + 2.0 * b * bm - 2.0 * cl), hfb = 2.0 * w * r * bm;
            if (lane < M) {
                gi += w * (2.0 * r * r * t + 2.0 * r * b * bm - 2.0 * r * cl);
#pragma unroll
                for (int q = 0; q < M; ++q) h[q] += 2.0 * w * r * r * brow[q];
                h[M + s] = hfr;
                h[M + S + s] = hfb;
            }
#pragma unroll
            for (int q = 0; q < M; ++q) {                  // rows of the scale / background variables of this specimen
                const double x1 = __shfl_sync(0xffffffffu, hfr, q), x2 = __shfl_sync(0xffffffffu, hfb, q);
                if (lane == M + s) h[q] = x1;
                if (lane == M + S + s) h[q] = x2;
            }
            if (lane == M + s) { gi = w * (2.0 * r * fBf + 2.0 * b * Bcf - 2.0 * cf); h[M + s] = 2.0 * w * fBf; h[M + S + s] = 2.0 * w * Bcf; }
            if (lane == M + S + s) { gi = w * (2.0 * r * Bcf + 2.0 * BMM * b - 2.0 * cM); h[M + s] = 2.0 * w * Bcf; h[M + S + s] = 2.0 * w * BMM; }
        }
        double hd = 0.0;
#pragma unroll
        for (int j = 0; j < D; ++j) hd = (lane == j) ? h[j] : hd;
        const bool at_bound = (vi <= lo_i && gi > 0.0);
        const bool free_i = lane < D && !fixed_bg && !at_bound && hd > 0.0;
        const unsigned mask = __ballot_sync(0xffffffffu, free_i);
        if (mask == 0u) break;
#pragma unroll
        for (int j = 0; j < D; ++j)
            if (!free_i || !((mask >> j) & 1u)) h[j] = (lane == j) ? 1.0 : 0.0;
        const double gfree = free_i ? gi : 0.0;
#ifdef PX_OPT_PROFILE
        if (lane == 0 && warp == 0) { sm.prof[3] += clock64() - tA; sm.prof[7] += 1; }
#endif
        // the first PX_OPT_NDAMP warps of the block each take the step with their own damping lam * 5^(warp - 1) -- the
        // retries of a serial Levenberg-Marquardt loop side by side on warps that would otherwise wait; the best point
        // wins.  (More warps than this make every warp slower: the solver lives on shuffles.)
        bool accepted = false;
        double rel = 0.0;
#pragma unroll 1
        for (int round = 0; round < 4 && !accepted; ++round) {      // a round nobody wins moves the whole ladder up
#ifdef PX_OPT_PROFILE
            long long tC = clock64();
            if (lane == 0 && warp == 0) sm.prof[6] += 1;
#endif
            double lam_w = lam * 0.2;
            for (int i = 0; i < warp; ++i) lam_w *= 5.0;
            double a[D], ut[D], invd[D];
#pragma unroll
            for (int j = 0; j < D; ++j) { a[j] = (lane == j) ? h[j] + (lam_w * fabs(h[j]) + 1e-300) : h[j]; ut[j] = 0.0; }
            bool ok = true;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const double djj = __shfl_sync(0xffffffffu, a[j], j);
                if (!(djj > 0.0)) ok = false;                 // warp-uniform
                const double inv = rsqrt(ok ? djj : 1.0);
                invd[j] = inv;
                const double lij = lane > j ? a[j] * inv : 0.0;
                a[j] = lij;
#pragma unroll
                for (int k = j + 1; k < D; ++k) {
                    const double lkj = __shfl_sync(0xffffffffu, lij, k);
                    a[k] = fma(-lij, lkj, a[k]);
                    if (lane == j) ut[k] = lkj;
                }
            }
#ifdef PX_OPT_PROFILE
            if (lane == 0 && warp == 0) sm.prof[4] += clock64() - tC;
            tC = clock64();
#endif
            // L z = -g, L^T x = z; the diagonal of L is 1 / invd
            double y = -gfree;
#pragma unroll
            for (int j = 0; j < D; ++j) {
                const double zj = __shfl_sync(0xffffffffu, y, j) * invd[j];
                if (lane == j) y = zj;
                else if (lane > j) y = fma(-a[j], zj, y);
            }
#pragma unroll
            for (int j = D - 1; j >= 0; --j) {
                const double xj = __shfl_sync(0xffffffffu, y, j) * invd[j];
                if (lane == j) y = xj;
                else if (lane < j) y = fma(-ut[j], xj, y);
            }
            const double vn = (free_i && ok) ? fmax(lo_i, vi + y) : vi;
            const double qn = ok ? opt_q_reg<M, S>(sm, vn, lane) : qv;
            if (lane == 0) sm.cand_q[warp] = (ok && qn < qv) ? qn : qv;
            if (lane < D) sm.cand_v[warp][lane] = vn;
            opt_team_sync();
            int best = -1;
            double qbest = qv;
#pragma unroll
            for (int w2 = 0; w2 < NW; ++w2) {
                const double qw = sm.cand_q[w2];
                if (qw < qbest) { qbest = qw; best = w2; }
            }
            if (best >= 0) {
                accepted = true;
                rel = (qv - qbest) / fmax(fabs(qv), 1e-300);
                if (lane < D) vi = sm.cand_v[best][lane];
                double lam_b = lam * 0.2;
                for (int i = 0; i < best; ++i) lam_b *= 5.0;
                lam = fmax(lam_b, 1e-12);
            } else {
                for (int i = 0; i < NW; ++i) lam *= 5.0;      // beyond every damping tried
            }
            opt_team_sync();                                  // cand_* are rewritten by the next step
#ifdef PX_OPT_PROFILE
            if (lane == 0 && warp == 0) sm.prof[5] += clock64() - tC;
#endif
        }
        if (!accepted || rel < sm.newton_tol) break;
    }
    if (warp == 0) {
        if (lane < D) sm.v[lane] = vi;
        if (lane == 0) sm.lam = fmin(lam, 1e-3);
    }
}

// The same step for very small problems (D = M + 2 S <= 5: one or two phases in one specimen, one phase in two) with the
// whole system in the registers of ONE lane: the team version spends its time in shuffles whatever D is (assembly and
// majoriser values are butterflies), here an iteration is a few hundred scalar operations, a chain bound by the latency
// of the fp64 pipe.  The four dampings of the ladder (lam * 5^(w - 1), w = 0 .. 3) run on four neighbouring lanes of
// warp 0, the best point wins, a round nobody wins moves the ladder up.  Called by warp 0.
template <int M, int S>
__device__ __forceinline__ double opt_q_scalar(const OptSmem &sm, const double (&v)[M + 2 * S]) {
    constexpr int MB = M + 1;
    double q = 0.0;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const double *B = sm.B[s], *c = sm.c[s];
        const double r = v[M + s], b = v[M + S + s];
        double fBf = 0.0, Bcf = 0.0, cf = 0.0;
#pragma unroll
        for (int p = 0; p < M; ++p) {
            double t = 0.0;
#pragma unroll
            for (int q2 = 0; q2 < M; ++q2) t = fma(B[p * MB + q2], v[q2], t);
            fBf = fma(v[p], t, fBf);
            Bcf = fma(B[p * MB + M], v[p], Bcf);
            cf = fma(c[p], v[p], cf);
        }
        q += sm.wgt[s] * (r * r * fBf + 2.0 * r * b * Bcf + B[M * MB + M] * b * b - 2.0 * r * cf - 2.0 * c[M] * b);
    }
    return q;
}

template <int M, int S>
__device__ __noinline__ void opt_inner_solve_scalar(OptSmem &sm, int lm_iters, int refine_bg, int lane) {
    constexpr int MB = M + 1, D = M + 2 * S;
    double v[D], lo[D];
#pragma unroll
    for (int i = 0; i < D; ++i) { v[i] = sm.v[i]; lo[i] = sm.lo[i]; }
    double lam = sm.lam;
    for (int it = 0; it < lm_iters; ++it) {
        double H[D][D], g[D], qv = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            g[i] = 0.0;
#pragma unroll
            for (int j = 0; j < D; ++j) H[i][j] = 0.0;
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const double *B = sm.B[s], *c = sm.c[s];
            const double w = sm.wgt[s], r = v[M + s], b = v[M + S + s];
            const int ir = M + s, ib = M + S + s;
            double fBf = 0.0, Bcf = 0.0, cf = 0.0;
#pragma unroll
            for (int p = 0; p < M; ++p) {
                double t = 0.0;
#pragma unroll
                for (int q = 0; q < M; ++q) {
                    t = fma(B[p * MB + q], v[q], t);
                    H[p][q] += 2.0 * w * r * r * B[p * MB + q];
                }
                const double bm = B[p * MB + M], cl = c[p];
                fBf = fma(v[p], t, fBf);
                Bcf = fma(bm, v[p], Bcf);
                cf = fma(cl, v[p], cf);
                g[p] += w * (2.0 * r * r * t + 2.0 * r * b * bm - 2.0 * r * cl);
                const double hfr = w * (4.0 * r * t + 2.0 * b * bm - 2.0 * cl), hfb = 2.0 * w * r * bm;
                H[p][ir] = hfr; H[ir][p] = hfr;
                H[p][ib] = hfb; H[ib][p] = hfb;
            }
            const double BMM = B[M * MB + M], cM = c[M];
            qv += w * (r * r * fBf + 2.0 * r * b * Bcf + BMM * b * b - 2.0 * r * cf - 2.0 * cM * b);
            g[ir] = w * (2.0 * r * fBf + 2.0 * b * Bcf - 2.0 * cf);
            g[ib] = w * (2.0 * r * Bcf + 2.0 * BMM * b - 2.0 * cM);
            H[ir][ir] = 2.0 * w * fBf;
            H[ir][ib] = 2.0 * w * Bcf; H[ib][ir] = 2.0 * w * Bcf;
            H[ib][ib] = 2.0 * w * BMM;
        }
        bool fr[D], any = false;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const bool fixed_bg = (!refine_bg && i >= M + S);
            fr[i] = !fixed_bg && !(v[i] <= lo[i] && g[i] > 0.0) && H[i][i] > 0.0;
            any = any || fr[i];
        }
        if (!any) break;
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = 0; j < D; ++j)
                if (!fr[i] || !fr[j]) H[i][j] = (i == j) ? 1.0 : 0.0;
        bool accepted = false;
        double rel = 0.0;
#pragma unroll 1
        for (int round = 0; round < 4 && !accepted; ++round) {
            // the four dampings of the ladder on four neighbouring lanes (the other lanes of the warp repeat them)
            const int wdx = lane & (PX_OPT_NDAMP - 1);
            double lam_w = lam * 0.2;
            for (int i = 0; i < wdx; ++i) lam_w *= 5.0;
            double A[D][D], y[D], vn[D];
            bool ok = true;
#pragma unroll
            for (int i = 0; i < D; ++i) {
#pragma unroll
                for (int j = 0; j < D; ++j) A[i][j] = H[i][j];
                A[i][i] = H[i][i] + (lam_w * fabs(H[i][i]) + 1e-300);
                y[i] = fr[i] ? -g[i] : 0.0;
            }
#pragma unroll
            for (int j = 0; j < D; ++j) {                     // Cholesky, lower triangle in place; A[j][j] <- 1 / l_jj
                double djj = A[j][j];
#pragma unroll
                for (int k = 0; k < j; ++k) djj = fma(-A[j][k], A[j][k], djj);
                if (!(djj > 0.0)) ok = false;
                const double inv = rsqrt(ok ? djj : 1.0);
#pragma unroll
                for (int i = j + 1; i < D; ++i) {
                    double a = A[i][j];
#pragma unroll
                    for (int k = 0; k < j; ++k) a = fma(-A[i][k], A[j][k], a);
                    A[i][j] = a * inv;
                }
                A[j][j] = inv;
            }
#pragma unroll
            for (int j = 0; j < D; ++j) {                     // L z = -g
                double z = y[j];
#pragma unroll
                for (int k = 0; k < j; ++k) z = fma(-A[j][k], y[k], z);
                y[j] = z * A[j][j];
            }
#pragma unroll
            for (int j = D - 1; j >= 0; --j) {                // L^T x = z
                double z = y[j];
#pragma unroll
                for (int k = j + 1; k < D; ++k) z = fma(-A[k][j], y[k], z);
                y[j] = z * A[j][j];
            }
#pragma unroll
            for (int i = 0; i < D; ++i) vn[i] = fr[i] ? fmax(lo[i], v[i] + y[i]) : v[i];
            const double qn = opt_q_scalar<M, S>(sm, vn);
            double qb = (ok && qn < qv) ? qn : __longlong_as_double(0x7ff0000000000000LL);        // best of the four: lowest value, then lowest damping
            int wb = wdx;
#pragma unroll
            for (int o = 1; o < PX_OPT_NDAMP; o <<= 1) {
                const double q2 = __shfl_xor_sync(0xffffffffu, qb, o);
                const int w2 = __shfl_xor_sync(0xffffffffu, wb, o);
                if (q2 < qb || (q2 == qb && w2 < wb)) { qb = q2; wb = w2; }
            }
            if (qb < qv) {
                accepted = true;
                rel = (qv - qb) / fmax(fabs(qv), 1e-300);
                const int src = (lane & ~(PX_OPT_NDAMP - 1)) | wb;
#pragma unroll
                for (int i = 0; i < D; ++i) v[i] = __shfl_sync(0xffffffffu, vn[i], src);
                lam = fmax(__shfl_sync(0xffffffffu, lam_w, src), 1e-12);
            } else {
                for (int i = 0; i < PX_OPT_NDAMP; ++i) lam *= 5.0;
            }
        }
        if (!accepted || rel < sm.newton_tol) break;
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < D; ++i) sm.v[i] = v[i];
        sm.lam = fmin(lam, 1e-3);
    }
}

// The same step for 5 < D <= 12 on single lanes of warp 0 (`optimizer_lane_solver` = number of dampings, each on its own
// lane): the team version above pays ~10 k cycles per Levenberg-Marquardt round for D = 10 -- chains of shuffles, a named
// barrier and a round trip through shared memory between four warps -- for ~600 floating-point operations.  Here a lane
// carries the whole system and uses its block structure: the scale and background of a specimen couple to the fractions
// but not to another specimen's pair, so the damped Newton matrix is
//     [ Hff  C_1  C_2 .. ]      Hff: M x M,  C_s = [u_s v_s]: M x 2,  D_s: 2 x 2
//     [ C_1' D_1         ]
//     [ C_2'      D_2    ]
// and the pairs are eliminated in closed form (2 x 2 inverses): A = Hff - sum_s C_s D_s^-1 C_s' is the only matrix that
// is factorised -- M instead of M + 2 S dependent pivots (reciprocal square roots), a triangle of M (M + 1) / 2 doubles
// that stays in registers.  Variables that are not free (fixed background, held at a bound by the gradient) get identity
// rows, exactly as in the other solvers; the matrix is positive definite iff every D_s and A are.  The Hessian blocks are
// assembled once per iteration into shared memory (every lane computes the same values, lane 0 stores) and read back
// with broadcast loads in the rounds.  First version of the lane solver (round 2, full D x D left-looking Cholesky with
// the factor in local memory): ~9 k cycles per round on c4; this one: see DESIGN.md section 4.
__device__ __forceinline__ double px_rcp_newton(double x) {        // 1 / x for x > 0: special-function seed + two Newton steps
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}

template <int M, int S, int ND>
__device__ __noinline__ void opt_inner_solve_lane(OptSmem &sm, int lm_iters, int refine_bg, int lane) {
    constexpr int MB = M + 1, D = M + 2 * S;
    static_assert(ND == 4 || ND == 8 || ND == 16, "dampings side by side");
    double v[D];
#pragma unroll
    for (int i = 0; i < D; ++i) v[i] = sm.v[i];
    double lam = sm.lam;
    double *H = sm.H;                                           // rows 0 .. M-1: lower triangle of Hff; row M + s: u_s; row M + S + s: v_s
    for (int it = 0; it < lm_iters; ++it) {
        double g[D], hrr[S], hrb[S], hbb[S], qv = 0.0;
#pragma unroll
        for (int i = 0; i < D; ++i) g[i] = 0.0;
        __syncwarp();                                           // the previous iteration's reads of H are done
        {
            double c2[S];                                       // 2 w r^2 per specimen
#pragma unroll
            for (int s = 0; s < S; ++s) c2[s] = 2.0 * sm.wgt[s] * v[M + s] * v[M + s];
#pragma unroll
            for (int p = 0; p < M; ++p)
#pragma unroll
                for (int q = 0; q <= p; ++q) {
                    double h = 0.0;
#pragma unroll
                    for (int s = 0; s < S; ++s) h = fma(c2[s], sm.B[s][p * MB + q], h);
                    if (lane == 0) H[p * PX_OPT_LDA + q] = h;
                }
        }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const double *B = sm.B[s], *c = sm.c[s];
            const double w = sm.wgt[s], r = v[M + s], b = v[M + S + s];
            const int ir = M + s, ib = M + S + s;
            double fBf = 0.0, Bcf = 0.0, cf = 0.0;
#pragma unroll
            for (int p = 0; p < M; ++p) {
                double t0 = 0.0, t1 = 0.0;                      // two partial sums: half the dependent chain
#pragma unroll
                for (int q = 0; q < M; q += 2) {
                    t0 = fma(B[p * MB + q], v[q], t0);
                    if (q + 1 < M) t1 = fma(B[p * MB + q + 1], v[q + 1], t1);
                }
                const double t = t0 + t1;
                const double bm = B[p * MB + M], cl = c[p];
                fBf = fma(v[p], t, fBf);
                Bcf = fma(bm, v[p], Bcf);
                cf = fma(cl, v[p], cf);
                g[p] += w * (2.0 * r * r * t + 2.0 * r * b * bm - 2.0 * r * cl);
                if (lane == 0) {
                    H[ir * PX_OPT_LDA + p] = w * (4.0 * r * t + 2.0 * b * bm - 2.0 * cl);
                    H[ib * PX_OPT_LDA + p] = 2.0 * w * r * bm;
                }
            }
            const double BMM = B[M * MB + M], cM = c[M];
            qv += w * (r * r * fBf + 2.0 * r * b * Bcf + BMM * b * b - 2.0 * r * cf - 2.0 * cM * b);
            g[ir] = w * (2.0 * r * fBf + 2.0 * b * Bcf - 2.0 * cf);
            g[ib] = w * (2.0 * r * Bcf + 2.0 * BMM * b - 2.0 * cM);
            hrr[s] = 2.0 * w * fBf;
            hrb[s] = 2.0 * w * Bcf;
            hbb[s] = 2.0 * w * BMM;
        }
        __syncwarp();
        unsigned fr = 0u;                                       // bit i: variable i is free
#pragma unroll
        for (int p = 0; p < M; ++p)
            if (!(v[p] <= sm.lo[p] && g[p] > 0.0) && H[p * PX_OPT_LDA + p] > 0.0) fr |= 1u << p;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            if (!(v[M + s] <= sm.lo[M + s] && g[M + s] > 0.0) && hrr[s] > 0.0) fr |= 1u << (M + s);
            if (refine_bg && !(v[M + S + s] <= sm.lo[M + S + s] && g[M + S + s] > 0.0) && hbb[s] > 0.0) fr |= 1u << (M + S + s);
        }
        if (fr == 0u) break;
        bool accepted = false;
        double rel = 0.0;
#pragma unroll 1
        for (int round = 0; round < 4 && !accepted; ++round) {
            const int wdx = lane & (ND - 1);
#ifdef PX_OPT_PROFILE
            if (lane == 0) sm.prof[6] += 1;
#endif
            double lam_w = lam * 0.2;
            for (int i = 0; i < wdx; ++i) lam_w *= 5.0;
            bool ok = true;
            // A = damped Hff (identity rows for fractions that are not free), rhs = -g_f
            double A[M * (M + 1) / 2], y[M];                    // packed lower triangle: (i, j) at i (i + 1) / 2 + j
#pragma unroll
            for (int p = 0; p < M; ++p) {
                const bool fp = (fr >> p) & 1u;
#pragma unroll
                for (int q = 0; q < p; ++q) A[p * (p + 1) / 2 + q] = (fp && ((fr >> q) & 1u)) ? H[p * PX_OPT_LDA + q] : 0.0;
                const double hpp = fp ? H[p * PX_OPT_LDA + p] : 1.0;
                A[p * (p + 1) / 2 + p] = hpp + (lam_w * fabs(hpp) + 1e-300);
                y[p] = fp ? -g[p] : 0.0;
            }
            // eliminate the (scale, background) pair of every specimen
            double i11[S], i12[S], i22[S], er[S], eb[S];        // D_s^-1 and D_s^-1 (g_r, g_b)
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const bool f_r = (fr >> (M + s)) & 1u, f_b = (fr >> (M + S + s)) & 1u;
                const double h1 = f_r ? hrr[s] : 1.0, h2 = f_b ? hbb[s] : 1.0;
                const double a = h1 + (lam_w * fabs(h1) + 1e-300), c = h2 + (lam_w * fabs(h2) + 1e-300);
                const double bq = (f_r && f_b) ? hrb[s] : 0.0;
                const double det = fma(a, c, -bq * bq);
                if (!(det > 0.0)) ok = false;
                const double idet = px_rcp_newton(ok ? det : 1.0);
                i11[s] = c * idet; i12[s] = -bq * idet; i22[s] = a * idet;
                const double gr = f_r ? g[M + s] : 0.0, gb = f_b ? g[M + S + s] : 0.0;
                er[s] = fma(i11[s], gr, i12[s] * gb);
                eb[s] = fma(i12[s], gr, i22[s] * gb);
                double wu[M], wv[M], uu[M], vv[M];
#pragma unroll
                for (int p = 0; p < M; ++p) {
                    const bool fp = (fr >> p) & 1u;
                    uu[p] = (fp && f_r) ? H[(M + s) * PX_OPT_LDA + p] : 0.0;
                    vv[p] = (fp && f_b) ? H[(M + S + s) * PX_OPT_LDA + p] : 0.0;
                    wu[p] = fma(uu[p], i11[s], vv[p] * i12[s]);
                    wv[p] = fma(uu[p], i12[s], vv[p] * i22[s]);
                    y[p] = fma(wu[p], gr, fma(wv[p], gb, y[p]));
                }
#pragma unroll
                for (int p = 0; p < M; ++p)
#pragma unroll
                    for (int q = 0; q <= p; ++q)
                        A[p * (p + 1) / 2 + q] = fma(-wu[p], uu[q], fma(-wv[p], vv[q], A[p * (p + 1) / 2 + q]));
            }
            // right-looking Cholesky of A in place (diagonal <- 1 / l_jj), forward substitution alongside
#pragma unroll
            for (int j = 0; j < M; ++j) {
                const double djj = A[j * (j + 1) / 2 + j];
                if (!(djj > 0.0)) ok = false;
                const double inv = rsqrt(ok ? djj : 1.0);
                A[j * (j + 1) / 2 + j] = inv;
                y[j] *= inv;
#pragma unroll
                for (int i = j + 1; i < M; ++i) A[i * (i + 1) / 2 + j] *= inv;
#pragma unroll
                for (int i = j + 1; i < M; ++i) {
                    const double lij = A[i * (i + 1) / 2 + j];
                    y[i] = fma(-lij, y[j], y[i]);
#pragma unroll
                    for (int k2 = j + 1; k2 <= i; ++k2) A[i * (i + 1) / 2 + k2] = fma(-lij, A[k2 * (k2 + 1) / 2 + j], A[i * (i + 1) / 2 + k2]);
                }
            }
#pragma unroll
            for (int j = M - 1; j >= 0; --j) {                  // L^T x = z
                double z = y[j];
#pragma unroll
                for (int k2 = j + 1; k2 < M; ++k2) z = fma(-A[k2 * (k2 + 1) / 2 + j], y[k2], z);
                y[j] = z * A[j * (j + 1) / 2 + j];
            }
            double vn[D];
#pragma unroll
            for (int p = 0; p < M; ++p) vn[p] = ((fr >> p) & 1u) ? fmax(sm.lo[p], v[p] + y[p]) : v[p];
#pragma unroll
            for (int s = 0; s < S; ++s) {                       // (dr, db) = -D^-1 (g_rb + C' df)
                const bool f_r = (fr >> (M + s)) & 1u, f_b = (fr >> (M + S + s)) & 1u;
                double cu = 0.0, cv = 0.0;
#pragma unroll
                for (int p = 0; p < M; ++p) {
                    const bool fp = (fr >> p) & 1u;
                    cu = fma((fp && f_r) ? H[(M + s) * PX_OPT_LDA + p] : 0.0, y[p], cu);
                    cv = fma((fp && f_b) ? H[(M + S + s) * PX_OPT_LDA + p] : 0.0, y[p], cv);
                }
                const double dr = -(er[s] + fma(i11[s], cu, i12[s] * cv)), db = -(eb[s] + fma(i12[s], cu, i22[s] * cv));
                vn[M + s] = f_r ? fmax(sm.lo[M + s], v[M + s] + dr) : v[M + s];
                vn[M + S + s] = f_b ? fmax(sm.lo[M + S + s], v[M + S + s] + db) : v[M + S + s];
            }
            const double qn = opt_q_scalar<M, S>(sm, vn);
            double qb = (ok && qn < qv) ? qn : __longlong_as_double(0x7ff0000000000000LL);        // best of the ladder: lowest value, then lowest damping
            int wb = wdx;
#pragma unroll
            for (int o = 1; o < ND; o <<= 1) {
                const double q2 = __shfl_xor_sync(0xffffffffu, qb, o);
                const int w2 = __shfl_xor_sync(0xffffffffu, wb, o);
                if (q2 < qb || (q2 == qb && w2 < wb)) { qb = q2; wb = w2; }
            }
            if (qb < qv) {
                accepted = true;
                rel = (qv - qb) / fmax(fabs(qv), 1e-300);
                const int src = (lane & ~(ND - 1)) | wb;
#pragma unroll
                for (int i = 0; i < D; ++i) v[i] = __shfl_sync(0xffffffffu, vn[i], src);
                lam = fmax(__shfl_sync(0xffffffffu, lam_w, src), 1e-12);
            } else {
                for (int i = 0; i < ND; ++i) lam *= 5.0;
            }
        }
        if (!accepted || rel < sm.newton_tol) break;
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < D; ++i) sm.v[i] = v[i];
        sm.lam = fmin(lam, 1e-3);
    }
}

// register version (PX_OPT_NDAMP warps of the block, one damping each) where the problem is small enough (S <= 3 and
// D = M + 2 S <= 12), shared-memory version (warp 0) otherwise.  Each kernel instantiation (M fixed) compiles at most
// three register solvers; D <= 5 goes to the single-thread scalar solver.  Called by every thread of the block.
#ifdef PX_OPT_NO_SCALAR
#define PX_OPT_SCALAR_D 0
#else
#define PX_OPT_SCALAR_D 5
#endif
template <int M>
__device__ __forceinline__ void opt_solve_dispatch(OptSmem &sm, int S, int lm_iters, int refine_bg, int tid) {
    const int lane = tid & 31, warp = tid >> 5;
#ifndef PX_OPT_NO_SCALAR
    if constexpr (M + 2 <= 5) { if (S == 1) { if (warp == 0) opt_inner_solve_scalar<M, 1>(sm, lm_iters, refine_bg, lane); return; } }
    if constexpr (M + 4 <= 5) { if (S == 2) { if (warp == 0) opt_inner_solve_scalar<M, 2>(sm, lm_iters, refine_bg, lane); return; } }
#endif
    if (sm.lane_solver) {                                       // block-uniform
        if constexpr (PX_OPT_SCALAR_D < M + 2 && M + 2 <= 12) { if (S == 1) { if (warp == 0) { if (sm.lane_solver >= 8) opt_inner_solve_lane<M, 1, 8>(sm, lm_iters, refine_bg, lane); else opt_inner_solve_lane<M, 1, 4>(sm, lm_iters, refine_bg, lane); } return; } }
        if constexpr (PX_OPT_SCALAR_D < M + 4 && M + 4 <= 12) { if (S == 2) { if (warp == 0) { if (sm.lane_solver >= 8) opt_inner_solve_lane<M, 2, 8>(sm, lm_iters, refine_bg, lane); else opt_inner_solve_lane<M, 2, 4>(sm, lm_iters, refine_bg, lane); } return; } }
    }
    const bool team = warp < PX_OPT_NDAMP;
    if constexpr (PX_OPT_SCALAR_D < M + 2 && M + 2 <= 12) { if (S == 1) { if (team) opt_inner_solve_reg<M, 1>(sm, lm_iters, refine_bg, lane, warp); return; } }
    if constexpr (PX_OPT_SCALAR_D < M + 4 && M + 4 <= 12) { if (S == 2) { if (team) opt_inner_solve_reg<M, 2>(sm, lm_iters, refine_bg, lane, warp); return; } }
    if constexpr (M + 6 <= 12) { if (S == 3) { if (team) opt_inner_solve_reg<M, 3>(sm, lm_iters, refine_bg, lane, warp); return; } }
    if (warp == 0) opt_inner_solve(sm, M, S, M + 1, lm_iters, refine_bg, lane);
}

// TPB = 256 for populations that fit the machine in one wave (shortest chain per candidate), 128 for larger ones
// (twice as many candidates in flight: the Newton steps, which keep four warps busy, overlap better)
// Warp totals of N accumulators with ~N instead of 5 N shuffles: at every butterfly step a lane keeps one half of its
// values and hands the other half to its partner (lanes with bit O set keep the upper half).  After the five steps a
// lane holds ceil(N / 32) totals; opt_store_reduced unwinds which entries they are and writes them to out[0 .. N).
template <int N, int COUNT, int O>
__device__ __forceinline__ void opt_warp_reduce_many(double (&acc)[N], int lane) {
    constexpr int HALF = (COUNT + 1) / 2;
    const bool upper = (lane & O) != 0;
#pragma unroll
    for (int i = 0; i < HALF; ++i) {
        const double a = acc[i];
        const double b = (i + HALF < COUNT) ? acc[i + HALF] : 0.0;
        const double send = upper ? a : b, keep = upper ? b : a;
        acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, O);
    }
    if constexpr (O > 1) opt_warp_reduce_many<N, HALF, O / 2>(acc, lane);
}

// level (COUNT, O) of the unwinding: `base` = offset accumulated by the finer levels is added AFTER this level's own
// offset is known, so recurse first to the finest level and add offsets on the way back
template <int N, int COUNT, int O>
__device__ __forceinline__ void opt_store_reduced(const double (&acc)[N], int lane, int, bool, double *out) {
    // sizes of the five levels
    constexpr int C0 = COUNT, H0 = (C0 + 1) / 2, H1 = (H0 + 1) / 2, H2 = (H1 + 1) / 2, H3 = (H2 + 1) / 2, H4 = (H3 + 1) / 2;
    static_assert(O == 16, "unwinds the five steps 16, 8, 4, 2, 1");
#pragma unroll
    for (int j = 0; j < H4; ++j) {
        // index inside the input of step 1 (size H3), then step 2 (H2), 4 (H1), 8 (H0), 16 (C0)
        int idx = j + ((lane & 1) ? H4 : 0);
        bool ok = idx < H3;
        idx += (lane & 2) ? H3 : 0;
        ok = ok && idx < H2;
        idx += (lane & 4) ? H2 : 0;
        ok = ok && idx < H1;
        idx += (lane & 8) ? H1 : 0;
        ok = ok && idx < H0;
        idx += (lane & 16) ? H0 : 0;
        ok = ok && idx < C0;
        if (ok) out[idx] = acc[j];
    }
}

// Gram pass: register loads, PIF points in flight per thread.  Three shared-memory feeds were built and measured in round
// 2 and are NOT kept (numbers on c4, 256 candidates, B200): 8-byte cp.async per point in 256-point
// tiles (1.39 ms against 1.01 ms: twice the LDGSTS instructions for the same bytes); cp.async.bulk + mbarrier ring, one 2 KB
// row segment per copy (0.98 against 0.84 ms: a UBLKCP stalls the issuing warp ~136 cycles, 1090 cycles per 8-row tile
// whoever issues it, with one or two CTAs per SM); 16-byte cp.async of point pairs in 512-point tiles (0.83 against 0.79 ms,
// 0.96 with the two points of a pair interleaved: with 36 accumulators per thread at the 128-register cap the extra
// pointers and the second point spill).  Cycle counters show why none helps: the pass is not waiting for L2 (wait on the
// copies: 5 % of the pass), it is a chain of dependent instructions per warp at two to four warps per scheduler.
template <int MB, int TPB>
__global__ void __launch_bounds__(TPB, 512 / TPB) k_mixture_optimize(StaticDev st, BatchDev bt, OptParams op,
                                                                 double *solutions, double *opt_residuals) {
    constexpr int M = MB - 1;
    constexpr int NG = MB * (MB + 1) / 2;       // upper triangle of the Gram matrix
    constexpr int NACC = NG + MB + 1;           // + c + sum|r|
    __shared__ OptSmem sm;
    const int k = blockIdx.x, S = st.S, Z = st.Z, tid = threadIdx.x;
    const int d = M + 2 * S;
    if (tid == 0) {
#ifdef PX_OPT_PROFILE
        for (int i = 0; i < 8; ++i) sm.prof[i] = 0;
#endif
    }
#ifdef PX_OPT_PROFILE
    const long long t_begin = clock64();
#endif
    if (tid == 0) {
        sm.lam = 1e-3;
        sm.newton_tol = op.newton_tol;
        sm.lane_solver = op.lane_solver;
        for (int p = 0; p < M; ++p) { sm.v[p] = 1.0 / M; sm.lo[p] = 0.0; }             // mixture.py:60-68 start
        for (int s = 0; s < S; ++s) {
            sm.v[M + s] = 1.0; sm.lo[M + s] = 1e-12;
            sm.v[M + S + s] = op.refine_bg ? 0.0 : bt.bgshifts[k * S + s]; sm.lo[M + S + s] = 0.0;
            const double den = st.derived[s * 8 + 3], count = st.derived[s * 8 + 5];
            sm.wgt[s] = den > 0.0 ? 100.0 / (S * den) : 0.0;
            sm.mean_obs[s] = (den > 0.0 && count > 0.0) ? den / count : 1.0;
        }
    }
    __syncthreads();
    double delta_pow = 1.0;                                 // shrink^it
    for (int it = 0; it <= op.outer; ++it, delta_pow *= op.shrink) {
        const double delta_rel = fmax(op.delta0 * delta_pow, op.delta_min);
        for (int s = 0; s < S; ++s) {
            const int X = st.spec_npts[s];
            const long long soff = st.spec_off[s];
            const double *I = bt.intensities + (long long)k * bt.cand_stride + (long long)M * Z * soff;
            const double *obs = st.observed + (long long)Z * soff;
            const unsigned char *sel = st.selected + soff;
            const double *corr = st.corr + soff;
            double theta[MB];
#pragma unroll
            for (int p = 0; p < M; ++p) theta[p] = sm.v[M + s] * sm.v[p];
            theta[M] = sm.v[M + S + s];
            double acc[NACC];
#pragma unroll
            for (int a = 0; a < NACC; ++a) acc[a] = 0.0;
            const double delta = delta_rel * sm.mean_obs[s];
            // coarse passes (optimizer_coarse_until / _stride): while the weight floor is still large a majoriser built
            // from every stride-th group of points -- another group every pass -- moves the iterate as well as the full one
            const int stride = (it < op.coarse_until && op.coarse_stride > 1) ? op.coarse_stride : 1;
            const int g0 = stride > 1 ? it % stride : 0;
#ifdef PX_OPT_PROFILE
            long long tG = clock64();
#endif
            // PIF points in flight per thread: the loads of all of them are issued before the arithmetic (few phases
            // leave registers for more points: the pass is bound by the latency of these loads)
            constexpr int PIF = MB <= 2 ? 8 : MB <= 4 ? 4 : 2;
            auto one_point = [&](const double (&phi)[MB], const double o) {
                double calc = 0.0;
#pragma unroll
                for (int p = 0; p < MB; ++p) calc = fma(theta[p], phi[p], calc);
                const double r = o - calc;
                // weight of the majoriser 1 / max(|r|, delta): the 20-bit reciprocal of the special-function unit
                // is enough (any positive weight majorises, the objective itself is summed exactly) and keeps
                // the division's Newton steps off the fp64 pipe
                double w;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(w) : "d"(fmax(fabs(r), delta)));
                acc[NACC - 1] += fabs(r);
                int e = 0;
#pragma unroll
                for (int p = 0; p < MB; ++p) {
                    const double wp = w * phi[p];
                    acc[NG + p] = fma(wp, o, acc[NG + p]);
#pragma unroll
                    for (int q = p; q < MB; ++q, ++e) acc[e] = fma(wp, phi[q], acc[e]);
                }
            };
            auto accumulate = [&](const double (&phi)[PIF][MB], const double (&o)[PIF], const bool (&on)[PIF]) {
#pragma unroll
                for (int u = 0; u < PIF; ++u)
                    if (on[u]) one_point(phi[u], o[u]);
            };
            if constexpr (MB <= PX_OPT_PTR_MB) {
                // few phases: the pass is bound by instruction issue, not by the fp64 pipe.  Full groups of PIF points
                // read at constant offsets from one running pointer per array (no index arithmetic per load); the
                // ragged end of the pattern goes through the same body with clamped offsets.
                auto group = [&](auto full_c, const double *Ix, const double *cx, const double *ox,
                                 const unsigned char *sx, int left) {
                    constexpr bool FULL = decltype(full_c)::value;
                    double phi[PIF][MB], o[PIF];
                    bool on[PIF];
#pragma unroll
                    for (int u = 0; u < PIF; ++u) {
                        const int off = FULL ? u * TPB : (u * TPB < left ? u * TPB : 0);
                        on[u] = (FULL || u * TPB < left) && sx[off];
#pragma unroll
                        for (int p = 0; p < M; ++p) phi[u][p] = Ix[(long long)p * Z * X + off];
                        phi[u][M] = cx[off];
                        o[u] = ox[off];
                    }
                    accumulate(phi, o, on);
                };
                for (int z = 0; z < Z; ++z) {
                    const double *Ix = I + (long long)z * X + tid, *ox = obs + (long long)z * X + tid;
                    const double *cx = corr + tid;
                    const unsigned char *sx = sel + tid;
                    // (coarse passes: every `stride`-th group of PIF * TPB points, starting at group `g0`)
                    const int skip = g0 * PIF * TPB, hop = PIF * TPB * stride;
                    int left = X - tid - skip;              // points from this thread's position to the end
                    Ix += skip; ox += skip; cx += skip; sx += skip;
                    for (; left > (PIF - 1) * TPB; left -= hop, Ix += hop, ox += hop, cx += hop, sx += hop)
                        group(std::true_type(), Ix, cx, ox, sx, left);
                    if (left > 0) group(std::false_type(), Ix, cx, ox, sx, left);
                }
            } else {
                // many phases: indexed loads (a running pointer per phase would cost the registers of the accumulators)
                for (int z = 0; z < Z; ++z)
                    for (int x0 = tid + g0 * PIF * TPB; x0 < X; x0 += PIF * TPB * stride) {
                        double phi[PIF][MB], o[PIF];
                        bool on[PIF];
#pragma unroll
                        for (int u = 0; u < PIF; ++u) {
                            const int x = x0 + u * TPB;
                            on[u] = x < X && sel[x];
                            const int xc = x < X ? x : x0;
#pragma unroll
                            for (int p = 0; p < M; ++p) phi[u][p] = I[((long long)p * Z + z) * X + xc];
                            phi[u][M] = corr[xc];
                            o[u] = obs[(long long)z * X + xc];
                        }
                        accumulate(phi, o, on);
                    }
            }
            // block reduction of NACC (+ count) values
            const int lane = tid & 31, warp = tid >> 5;
#ifdef PX_OPT_PROFILE
            if (tid == 0) sm.prof[0] += clock64() - tG;
            tG = clock64();
#endif
            opt_warp_reduce_many<NACC, NACC, 16>(acc, lane);
            opt_store_reduced<NACC, NACC, 16>(acc, lane, 0, true, sm.red[warp]);
            __syncthreads();
            if (tid < NACC) {
                double vsum = 0.0;
                for (int w2 = 0; w2 < TPB / 32; ++w2) vsum += sm.red[w2][tid];
                if (tid < NG) {
                    // unpack upper-triangle index tid -> (p, q) and mirror
                    int p = 0, e = tid;
                    while (e >= MB - p) { e -= MB - p; ++p; }
                    const int q = p + e;
                    sm.B[s][p * MB + q] = vsum;
                    sm.B[s][q * MB + p] = vsum;
                } else if (tid < NG + MB) {
                    sm.c[s][tid - NG] = vsum;
                } else {
                    sm.obj[s] = vsum;
                }
            }
            __syncthreads();
#ifdef PX_OPT_PROFILE
            if (tid == 0) sm.prof[1] += clock64() - tG;
#endif
        }
        if (it == op.outer) break;                 // the last pass only evaluates the objective
#ifdef PX_OPT_PROFILE
        long long tN = clock64();
#endif
        opt_solve_dispatch<M>(sm, S, (op.single_from > 0 && it >= op.single_from) ? 1 : op.lm_iters, op.refine_bg, tid);
        __syncthreads();
#ifdef PX_OPT_PROFILE
        if (tid == 0) sm.prof[2] += clock64() - tN;
#endif
    }
#ifdef PX_OPT_PROFILE
    if (tid == 0)
        printf("opt profile k=%d sm=%d total %lld: gram %lld reduce %lld newton %lld (assembly %lld chol %lld solve+q %lld) tries %lld lm %lld\n", k, (int)__smid(), clock64() - t_begin,
               sm.prof[0], sm.prof[1], sm.prof[2], sm.prof[3], sm.prof[4], sm.prof[5], sm.prof[6], sm.prof[7]);
#endif
    if (tid == 0) {
        // reference units: fractions = f / sum f, scales = rho * sum f (mixture.py:196-205)
        double fsum = 0.0;
        for (int p = 0; p < M; ++p) fsum += sm.v[p];
        double *out = solutions + (long long)k * d;
        for (int p = 0; p < M; ++p) out[p] = fsum > 0.0 ? sm.v[p] / fsum : 1.0 / M;
        for (int s = 0; s < S; ++s) {
            out[M + s] = sm.v[M + s] * (fsum > 0.0 ? fsum : 1.0);
            out[M + S + s] = sm.v[M + S + s];
        }
        double mean = 0.0;
        for (int s = 0; s < S; ++s) {
            const double den = st.derived[s * 8 + 3];
            const double rp = den > 0.0 ? sm.obj[s] / den * 100.0 : 0.0;
            opt_residuals[(long long)k * (S + 1) + 1 + s] = rp;
            mean += rp;
        }
        opt_residuals[(long long)k * (S + 1)] = mean / S;
    }
}
