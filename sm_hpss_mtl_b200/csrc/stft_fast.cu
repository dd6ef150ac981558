// K1, specialised: fused framing + window + two-pass register FFT + |X| for the (n_fft, hop) pairs the
// reference and BASELINE.json use (400/160, 512/160, 512/128, 1024/256, 2048/512).
//
// Replaces librosa.core.stft(center=False) + np.abs (lib/preprocessing.py:381,387,407,417,429,439).
// Same tiling as the generic kernel in stft.cu (one CTA = up to 16 consecutive frames of one clip, each sample
// read from HBM once), but every size is a compile-time constant:
//   * real FFT of size n_fft = complex FFT of size N2 = n_fft/2 on (even, odd) sample pairs;
//   * N2 = NA x NB.  Pass 1: for each b < NB a register DFT of size NA over the points NB*q + b (window multiply
//     fused into the loads); pass 2: for each k1 < NA the twiddles W_N2^(b*k1) and a register DFT of size NB
//     over b, written back in place -- which leaves the spectrum in natural order, so there are only two
//     shared-memory round trips (the generic kernel makes one per radix) and no index arithmetic at run time;
//   * the DFTs are generated straight-line codelets (tools/gen_fft_codelets.py), twiddles are literals;
//   * lane = frame everywhere (half-warps of 16 frames): the frame stride of the spectrum buffer (N2 + 1
//     float2) and of the padded sample staging ((hop + pad)/2 float2, pad chosen so that it is odd) make every
//     shared-memory access conflict free, and the magnitudes leave as 64-byte runs of one frequency row.
#include <stdlib.h>

#include "common.cuh"
#include "fft_codelets_gen.cuh"

namespace hpss {

namespace {

__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Batches of equal clips whose length is a multiple of the hop: frame t of clip c starts at sample
// (c * fpc + t) * hop, so the batch is one long signal in which the "virtual" frames t >= T of every clip are
// skipped; tiles then hold 16 consecutive virtual frames regardless of clip borders (T = 98: 98 % of the lanes
// busy instead of 14 of 16).  fpc = 0: per-clip tile list.
struct UniformBatch {
    int fpc;                 // virtual frames per clip = clip samples / hop
    int T;                   // real frames per clip
    int n_clips;
    int64_t total_samples;
};

template <int NFFT, int HOP, int NA, int NB, int NT>
struct FastCfg {
    static constexpr int N2 = NFFT / 2;
#ifndef HPSS_K1_TT400
#define HPSS_K1_TT400 16            // frames per tile of the 400/160 configuration (16 or 32; 32 measured 5 % slower)
#endif
#ifndef HPSS_K1_TT_BIG
#define HPSS_K1_TT_BIG 8            // frames per tile at n_fft = 2048: half the shared memory, two CTAs per SM instead of
#endif                              // one (0.516 -> 0.479 ms on 64 x 60 s; at n_fft = 1024 8-frame tiles are slower: 0.38 -> 0.49 ms)
    static constexpr int TT = (NFFT == 400) ? HPSS_K1_TT400 : (NFFT >= 2048 ? HPSS_K1_TT_BIG : 16);   // frames per tile = lanes per group
    static constexpr int ZS = N2 | 1;                                   // float2 stride between frames
    static constexpr int PAD = ((HOP / 2) % 2 == 0) ? 2 : 0;            // (HOP + PAD) / 2 odd
    static constexpr int HOPP = HOP + PAD;
    static constexpr int SEG = (TT - 1) * HOP + NFFT;                   // samples of a full tile
    static constexpr int SEGP = SEG + PAD * ((SEG + HOP - 1) / HOP);    // padded
    static constexpr int G = NT / TT;                                   // lane groups (one frame per lane) per CTA
#ifndef HPSS_K1_GLOBAL_TABLES
#define HPSS_K1_GLOBAL_TABLES 1     // 1: window / twiddles are read through L1 (__ldg), not staged per CTA
#endif
    static constexpr bool GT = HPSS_K1_GLOBAL_TABLES != 0;
    static constexpr size_t smem_bytes =
        sizeof(float2) * ((size_t)TT * ZS + (GT ? 0 : N2 + (N2 / 2 + 1))) + sizeof(float) * ((size_t)(GT ? 0 : NFFT) + SEGP + 4);
    static_assert(NA * NB == N2, "N2 = NA * NB");
    static_assert(HOP % (2 * NB) == 0, "pad offsets must be compile-time constants");
    static_assert(HOP % 4 == 0 && NT % 32 == 0, "vector staging");
};

template <int NFFT, int HOP, int NA, int NB, int NT, int MODE>
__global__ void __launch_bounds__(NT)
stft_fast_kernel(const float* __restrict__ wave, const int64_t* __restrict__ sample_off,
                 const int64_t* __restrict__ frame_off, const int2* __restrict__ tiles,
                 const float* __restrict__ window, const float2* __restrict__ tw_half,
                 const float2* __restrict__ tw_full, const float2* __restrict__ win_bq,
                 const float2* __restrict__ tw_kb, int power, float* __restrict__ S, float2* __restrict__ cplx,
                 UniformBatch uni) {
    using C = FastCfg<NFFT, HOP, NA, NB, NT>;
    constexpr int N2 = C::N2, ZS = C::ZS, PAD = C::PAD, HOPP = C::HOPP, G = C::G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* Z = reinterpret_cast<float2*>(smem_raw);            // [TT][ZS]
    float2* s_twh_ = Z + (size_t)C::TT * ZS;                    // [N2]      exp(-2 pi i k / N2)
    float2* s_twf_ = s_twh_ + (C::GT ? 0 : N2);                 // [N2/2+1]  exp(-2 pi i k / NFFT)
    float* s_win_ = reinterpret_cast<float*>(s_twf_ + (C::GT ? 0 : (N2 / 2 + 1)));
    float* s_samp = s_win_ + (C::GT ? 0 : NFFT);                // padded: sample s at s + PAD * (s / HOP)
    // tables: staged copies, or (GT) the L1-resident global arrays (the window is then pre-halved on the host)
    const float2* s_twh = C::GT ? tw_half : s_twh_;
    const float2* s_twf = C::GT ? tw_full : s_twf_;
    const float* s_win = C::GT ? window : s_win_;

    const int tid = threadIdx.x;
    const int fr = tid % C::TT;
    const int g = tid / C::TT;
    int T, seg;
    int64_t base;                  // element offset of (bin 0, this lane's frame) in S
    bool live;
    const float* src;
    if (uni.fpc > 0) {
        const int64_t v0 = (int64_t)blockIdx.x * C::TT;
        src = wave + v0 * HOP;
        const int64_t left = uni.total_samples - v0 * HOP;
        seg = (int)(left < (int64_t)C::SEG ? left : (int64_t)C::SEG);
        const uint32_t v = (uint32_t)v0 + fr;              // virtual frames fit 32 bits (checked at launch)
        const int c = (int)(v / (uint32_t)uni.fpc);
        const int t = (int)(v - (uint32_t)c * (uint32_t)uni.fpc);
        T = uni.T;
        live = c < uni.n_clips && t < T;
        base = (int64_t)(N2 + 1) * ((int64_t)c * T) + t;
    } else {
        const int2 tile = tiles[blockIdx.x];
        const int c = tile.x, t0 = tile.y;
        const int64_t fo = frame_off[c];
        T = (int)(frame_off[c + 1] - fo);
        const int nf = min(C::TT, T - t0);
        src = wave + sample_off[c] + (int64_t)t0 * HOP;
        seg = (nf - 1) * HOP + NFFT;
        live = fr < nf;
        base = (int64_t)(N2 + 1) * fo + t0 + fr;
    }

    // ---- stage tables and the sample segment
    // window x 0.5 (exact): the spectrum buffer then holds Z/2 and the unpack needs no halving
    if (!C::GT) {
        for (int i = tid; i < NFFT / 2; i += NT) {
            const float2 wv = __ldg(reinterpret_cast<const float2*>(window) + i);
            reinterpret_cast<float2*>(s_win_)[i] = make_float2(0.5f * wv.x, 0.5f * wv.y);
        }
        for (int i = tid; i < N2; i += NT) s_twh_[i] = __ldg(tw_half + i);
        for (int i = tid; i <= N2 / 2; i += NT) s_twf_[i] = __ldg(tw_full + i);
    }
    {
        // 8-byte granularity: consecutive lanes write consecutive float2 (no bank conflicts; the padded layout
        // keeps 8-byte but not 16-byte alignment)
        if ((reinterpret_cast<uintptr_t>(src) & 7) == 0 && (seg & 1) == 0) {
            const float2* src2 = reinterpret_cast<const float2*>(src);
            for (int i = tid; i < seg / 2; i += NT) {
                const int s = 2 * i;
                *reinterpret_cast<float2*>(s_samp + s + PAD * (s / HOP)) = __ldg(src2 + i);
            }
        } else {
            for (int s = tid; s < seg; s += NT) s_samp[s + PAD * (s / HOP)] = __ldg(src + s);
        }
    }
    __syncthreads();

    // ---- pass 1: DFT-NA over q of w[n] x[n], n = NB*q + b  ->  Y[b][k1] at NA*b + k1
    if (live) {
        const float2* xs = reinterpret_cast<const float2*>(s_samp + fr * HOPP);
        const float2* ws = reinterpret_cast<const float2*>(s_win);
#pragma unroll 1
        for (int b = g; b < NB; b += G) {
            float2 v[NA];
            const float4* wq = reinterpret_cast<const float4*>(win_bq + b * NA);     // two q per 16-byte load
#pragma unroll
            for (int q = 0; q < NA; ++q) {
                const int j0 = NB * q;                                   // compile-time after unrolling
                const int po = (2 * j0 + PAD * ((2 * j0) / HOP)) / 2;    // float2 offset in the padded staging
                const float2 xv = xs[po + b];
                float2 wv;
                if (C::GT) {
                    static_assert(NA % 2 == 0, "window pairs");
                    const float4 w2 = __ldg(wq + q / 2);                 // CSE'd between q and q + 1
                    wv = (q & 1) ? make_float2(w2.z, w2.w) : make_float2(w2.x, w2.y);
                } else {
                    wv = ws[j0 + b];
                }
                v[q] = make_float2(xv.x * wv.x, xv.y * wv.y);
            }
            Dft<NA>::run(v);
            float2* z = Z + fr * ZS + NA * b;
#pragma unroll
            for (int k1 = 0; k1 < NA; ++k1) z[k1] = v[k1];
        }
    }
    __syncthreads();

    // ---- pass 2: twiddle W_N2^(b*k1), DFT-NB over b  ->  Z[k1 + NA*k2], in place (natural order)
    if (live) {
#pragma unroll 1
        for (int k1 = g; k1 < NA; k1 += G) {
            float2 v[NB];
            float2* z = Z + fr * ZS + k1;
            v[0] = z[0];
            const float4* tq = reinterpret_cast<const float4*>(tw_kb + k1 * NB);     // two b per 16-byte load
#pragma unroll
            for (int b = 1; b < NB; ++b) {
                const float2 a = z[NA * b];
                float2 w;
                if (C::GT) {
                    static_assert(NB % 2 == 0, "twiddle pairs");
                    const float4 w2 = __ldg(tq + b / 2);
                    w = (b & 1) ? make_float2(w2.z, w2.w) : make_float2(w2.x, w2.y);
                } else {
                    w = s_twh[b * k1];
                }
                v[b] = make_float2(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x);
            }
            Dft<NB>::run(v);
#pragma unroll
            for (int k2 = 0; k2 < NB; ++k2) z[NA * k2] = v[k2];
            if (k1 == 0) z[N2] = v[0];          // Z[N2] = Z[0] (periodicity): the unpack reads Z[N2 - k] for k = 0 too
        }
    }
    __syncthreads();

    // ---- real-FFT unpack + magnitude, written (f, t) with lanes along t.  Bins k and N2-k share their loads:
    // X[k] = E + W*O and X[N2-k] = conj(E - W*O) with E, O from Z[k] and conj(Z[N2-k]).
    if (live) {
        const float2* zrow = Z + fr * ZS;
        char* Sg = reinterpret_cast<char*>(S + base);
        const int T4 = 4 * T;                              // row pitch in bytes: one IMAD.WIDE per address
        // one bin pair: X[k] = E + W*O, X[N2-k] = conj(E - W*O), E = Z[k] + conj(Z[N2-k]), O = -i (Z[k] - conj(Z[N2-k]))
        auto pair = [&](int k, float2& xa, float2& xb) {
            const float2 zk = zrow[k];
            const float2 zc = zrow[N2 - k];
            const float2 e = make_float2(zk.x + zc.x, zk.y - zc.y);
            const float2 o = make_float2(zk.y + zc.y, zc.x - zk.x);
            const float2 w = C::GT ? __ldg(s_twf + k) : s_twf[k];
            const float2 wo = make_float2(w.x * o.x - w.y * o.y, w.x * o.y + w.y * o.x);
            xa = make_float2(e.x + wo.x, e.y + wo.y);
            xb = make_float2(e.x - wo.x, wo.y - e.y);
        };
        if (MODE == 0) {                                   // magnitudes only (the feature path)
            static_assert((N2 / 2) % G == 0, "every group unpacks the same number of bin pairs");
#pragma unroll
            for (int i = 0; i < (N2 / 2) / G; ++i) {       // fully unrolled: table and buffer offsets become immediates
                const int k = g + i * G;
                float2 xa, xb;
                pair(k, xa, xb);
                *reinterpret_cast<float*>(Sg + (int64_t)k * T4) = fast_sqrt(xa.x * xa.x + xa.y * xa.y);
                *reinterpret_cast<float*>(Sg + (int64_t)(N2 - k) * T4) = fast_sqrt(xb.x * xb.x + xb.y * xb.y);
            }
            if (g == (N2 / 2) % G) {                       // the middle bin pairs with itself
                float2 xa, xb;
                pair(N2 / 2, xa, xb);
                *reinterpret_cast<float*>(Sg + (int64_t)(N2 / 2) * T4) = fast_sqrt(xa.x * xa.x + xa.y * xa.y);
            }
        } else {
            float2* Cg = cplx ? cplx + base : nullptr;
            for (int k = g; k <= N2 / 2; k += G) {
                const int k2 = N2 - k;
                float2 xa, xb;
                pair(k, xa, xb);
                const float pa = xa.x * xa.x + xa.y * xa.y;
                const float pb = xb.x * xb.x + xb.y * xb.y;
                const int64_t ga = (int64_t)k * T, gb = (int64_t)k2 * T;
                *reinterpret_cast<float*>(Sg + 4 * ga) = power ? pa : fast_sqrt(pa);
                if (Cg) Cg[ga] = xa;
                if (k2 != k) {
                    *reinterpret_cast<float*>(Sg + 4 * gb) = power ? pb : fast_sqrt(pb);
                    if (Cg) Cg[gb] = xb;
                }
            }
        }
    }
}


// ---- n_fft = 400, hop = 160, magnitudes only: real-input 20 x 20 split without an unpack step.
// X[k1 + 20 k2] = sum_b ( W_400^(b k1) * Y[b][k1] ) W_20^(b k2),  Y[b][k1] = sum_q w[n] x[n] W_20^(q k1), n = 20 q + b.
// Pass 1: twenty real-input DFT-20 codelets per frame (11 non-redundant outputs each; Y[b][0] and Y[b][10] are real
// and share one float2 slot).  Pass 2: for k1 = 0..10 the twiddles and a complex DFT-20 whose outputs are final bins
// (k <= 200) or mirrors of final bins (|X[400 - k]| = |X[k]|), so the magnitudes go from registers to global memory:
// 440 instead of 800 complex values per frame pass through shared memory, one barrier and the unpack loop are gone.
constexpr int kR400Threads = 192;         // 12 half-warp groups: 20 b over 12 groups in pass 1, 11 k1 in pass 2

__global__ void __launch_bounds__(kR400Threads)
stft_r400_kernel(const float* __restrict__ wave, const int64_t* __restrict__ sample_off,
                 const int64_t* __restrict__ frame_off, const int2* __restrict__ tiles,
                 const float* __restrict__ win_rq, const float2* __restrict__ tw_r, float* __restrict__ S,
                 UniformBatch uni) {
    using C = FastCfg<400, 160, 10, 20, kR400Threads>;
    constexpr int NFFT = 400, HOP = 160, NT = kR400Threads, ZS = 201, PAD = C::PAD, HOPP = C::HOPP, G = NT / 16;
    static_assert(C::TT == 16 && PAD == 2, "layout");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* Z = reinterpret_cast<float2*>(smem_raw);            // [16][ZS]: slot k1 (0..9) x 20 b; slot 0 = (Y[b][0], Y[b][10])
    float* s_samp = reinterpret_cast<float*>(Z + (size_t)16 * ZS);

    const int tid = threadIdx.x;
    const int fr = tid & 15;
    const int g = tid >> 4;
    int T, seg;
    int64_t base;
    bool live;
    const float* src;
    if (uni.fpc > 0) {
        const int64_t v0 = (int64_t)blockIdx.x * 16;
        src = wave + v0 * HOP;
        const int64_t left = uni.total_samples - v0 * HOP;
        seg = (int)(left < (int64_t)C::SEG ? left : (int64_t)C::SEG);
        const uint32_t v = (uint32_t)v0 + fr;
        const int c = (int)(v / (uint32_t)uni.fpc);
        const int t = (int)(v - (uint32_t)c * (uint32_t)uni.fpc);
        T = uni.T;
        live = c < uni.n_clips && t < T;
        base = (int64_t)201 * ((int64_t)c * T) + t;
    } else {
        const int2 tile = tiles[blockIdx.x];
        const int c = tile.x, t0 = tile.y;
        const int64_t fo = frame_off[c];
        T = (int)(frame_off[c + 1] - fo);
        const int nf = min(16, T - t0);
        src = wave + sample_off[c] + (int64_t)t0 * HOP;
        seg = (nf - 1) * HOP + NFFT;
        live = fr < nf;
        base = (int64_t)201 * fo + t0 + fr;
    }
    if ((reinterpret_cast<uintptr_t>(src) & 7) == 0 && (seg & 1) == 0) {
        // 160 threads stage two hop rows (2 x 80 float2) per step: no division, all loads issued before the stores
        if (tid < 160) {
            constexpr int STEPS = (C::SEG / HOP + 2) / 2;                 // 9 steps cover rows 0 .. 17
            const int row0 = tid >= 80 ? 1 : 0, col = tid - 80 * row0;
            const float2* sp = reinterpret_cast<const float2*>(src) + row0 * 80 + col;
            float* dp = s_samp + row0 * HOPP + 2 * col;
            const int lim = seg / 2 - (row0 * 80 + col);                  // float2 units left from this thread's start
            float2 v[STEPS];
#pragma unroll
            for (int j = 0; j < STEPS; ++j) v[j] = (160 * j < lim) ? __ldg(sp + 160 * j) : make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < STEPS; ++j)
                if (160 * j < lim) *reinterpret_cast<float2*>(dp + 2 * HOPP * j) = v[j];
        }
    } else {
        for (int s = tid; s < seg; s += NT) s_samp[s + PAD * (s / HOP)] = __ldg(src + s);
    }
    __syncthreads();

    // ---- pass 1: real-input DFT-20 over q of w[n] x[n], n = 20 q + b
    if (live) {
        const float* xs = s_samp + fr * HOPP;
        float2* z = Z + fr * ZS;
#pragma unroll 1
        for (int b = g; b < 20; b += G) {
            const float4* wq = reinterpret_cast<const float4*>(win_rq + b * 20);
            float x[20];
#pragma unroll
            for (int q4 = 0; q4 < 5; ++q4) {
                const float4 w4 = __ldg(wq + q4);
                // n = 20 q + b < 160 for q < 8, < 320 for q < 16 (b < 20): the pad offset is a compile-time constant
                x[4 * q4 + 0] = xs[20 * (4 * q4 + 0) + PAD * ((4 * q4 + 0) / 8) + b] * w4.x;
                x[4 * q4 + 1] = xs[20 * (4 * q4 + 1) + PAD * ((4 * q4 + 1) / 8) + b] * w4.y;
                x[4 * q4 + 2] = xs[20 * (4 * q4 + 2) + PAD * ((4 * q4 + 2) / 8) + b] * w4.z;
                x[4 * q4 + 3] = xs[20 * (4 * q4 + 3) + PAD * ((4 * q4 + 3) / 8) + b] * w4.w;
            }
            float2 y[11];
            rdft20(x, y);
            z[b] = make_float2(y[0].x, y[10].x);
#pragma unroll
            for (int k1 = 1; k1 < 10; ++k1) z[20 * k1 + b] = y[k1];
        }
    }
    __syncthreads();

    // ---- pass 2: twiddle W_400^(b k1), DFT-20 over b, |.| straight to global memory.
    // Groups 0 and 1 (one warp) take k1 = 0 and k1 = 10, whose inputs are real (the .x / .y halves of slot 0) and
    // whose outputs beyond k2 = 10 / 9 repeat rows already covered; groups 2..10 take k1 = 1..9.
    if (live && g <= 10) {
        char* Sg = reinterpret_cast<char*>(S + base);
        const int T4 = 4 * T;
        float2 v[20];
        if (g < 2) {                                            // warp-uniform
            const int k1 = 10 * g;
            const float* zr = reinterpret_cast<const float*>(Z + fr * ZS) + g;
            const float4* tq = reinterpret_cast<const float4*>(tw_r + k1 * 20);
#pragma unroll
            for (int b2 = 0; b2 < 10; ++b2) {
                const float4 w4 = __ldg(tq + b2);
                const float r0 = zr[2 * (2 * b2)], r1 = zr[2 * (2 * b2 + 1)];
                v[2 * b2] = make_float2(r0 * w4.x, r0 * w4.y);
                v[2 * b2 + 1] = make_float2(r1 * w4.z, r1 * w4.w);
            }
            Dft<20>::run(v);
            const int kmax = 10 - g;
#pragma unroll
            for (int k2 = 0; k2 <= 10; ++k2)
                if (k2 <= kmax)
                    *reinterpret_cast<float*>(Sg + (int64_t)(k1 + 20 * k2) * T4) =
                        fast_sqrt(v[k2].x * v[k2].x + v[k2].y * v[k2].y);
        } else {
            const int k1 = g - 1;
            const float2* zp = Z + fr * ZS + 20 * k1;
            const float4* tq = reinterpret_cast<const float4*>(tw_r + k1 * 20);
#pragma unroll
            for (int b2 = 0; b2 < 10; ++b2) {
                const float4 w4 = __ldg(tq + b2);
                const float2 a0 = zp[2 * b2], a1 = zp[2 * b2 + 1];
                v[2 * b2] = (b2 == 0) ? a0 : make_float2(a0.x * w4.x - a0.y * w4.y, a0.x * w4.y + a0.y * w4.x);
                v[2 * b2 + 1] = make_float2(a1.x * w4.z - a1.y * w4.w, a1.x * w4.w + a1.y * w4.z);
            }
            Dft<20>::run(v);
#pragma unroll
            for (int k2 = 0; k2 < 20; ++k2) {
                const int row = (k2 < 10) ? k1 + 20 * k2 : 20 * (20 - k2) - k1;      // bin, or its mirror 400 - bin
                *reinterpret_cast<float*>(Sg + (int64_t)row * T4) = fast_sqrt(v[k2].x * v[k2].x + v[k2].y * v[k2].y);
            }
        }
    }
}

int launch_r400(hpss_ctx* ctx, hpss_batch* b, const float* wave, const FftPlan* plan, float* S, cudaStream_t st) {
    using C = FastCfg<400, 160, 10, 20, kR400Threads>;
    constexpr size_t smem = sizeof(float2) * 16 * 201 + sizeof(float) * (C::SEGP + 4);
    HPSS_CUDA(cudaFuncSetAttribute(stft_r400_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    UniformBatch uni{0, 0, 0, 0};
    const int64_t L = b->uniform_samples;
    if (L > 0 && L % 160 == 0 && b->uniform_frames > 0 && (int64_t)b->n_clips * (L / 160) + 16 < 0x7fffffff &&
        !knobs().no_uniform_stft) {
        uni.fpc = (int)(L / 160);
        uni.T = (int)b->uniform_frames;
        uni.n_clips = b->n_clips;
        uni.total_samples = b->sample_off[b->n_clips];
        const int64_t n_tiles = ((int64_t)b->n_clips * uni.fpc + 15) / 16;
        stft_r400_kernel<<<(unsigned)n_tiles, kR400Threads, smem, st>>>(wave, b->d_sample_off, b->d_frame_off, nullptr,
                                                                        plan->d_win_r400, plan->d_tw_r400, S, uni);
    } else {
        int tt = 16;
        if (b->max_frames > 0) {
            const int64_t nt = (b->max_frames + tt - 1) / tt;
            tt = (int)((b->max_frames + nt - 1) / nt);
        }
        int rc = ensure_stft_tiles(b, tt);
        if (rc) return rc;
        if (b->n_stft_tiles == 0) return HPSS_OK;
        stft_r400_kernel<<<b->n_stft_tiles, kR400Threads, smem, st>>>(wave, b->d_sample_off, b->d_frame_off,
                                                                      b->d_stft_tiles, plan->d_win_r400,
                                                                      plan->d_tw_r400, S, uni);
    }
    HPSS_LAUNCHED("stft_r400_kernel");
    return HPSS_OK;
}

template <int NFFT, int HOP, int NA, int NB, int NT>
int launch_cfg(hpss_ctx* ctx, hpss_batch* b, const float* wave, const FftPlan* plan, int power, float* S, float* cplx,
               cudaStream_t st) {
    using C = FastCfg<NFFT, HOP, NA, NB, NT>;
    if (C::smem_bytes > (size_t)ctx->max_smem_optin) {
        set_error("n_fft=%d does not fit the shared-memory FFT (max %d bytes)", NFFT, ctx->max_smem_optin);
        return HPSS_ERR_UNSUPPORTED;
    }
    // MODE 0: magnitudes only (the feature path); MODE 1: power and / or complex output as well
    auto kern = (power || cplx) ? stft_fast_kernel<NFFT, HOP, NA, NB, NT, 1> : stft_fast_kernel<NFFT, HOP, NA, NB, NT, 0>;
    HPSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes));
    UniformBatch uni{0, 0, 0, 0};
    const int64_t L = b->uniform_samples;
    if (L > 0 && L % HOP == 0 && b->uniform_frames > 0 && (int64_t)b->n_clips * (L / HOP) + C::TT < 0x7fffffff && !knobs().no_uniform_stft) {
        // equal clips, length a multiple of the hop: 16 consecutive virtual frames per tile, across clip borders
        uni.fpc = (int)(L / HOP);
        uni.T = (int)b->uniform_frames;
        uni.n_clips = b->n_clips;
        uni.total_samples = b->sample_off[b->n_clips];
        const int64_t n_tiles = ((int64_t)b->n_clips * uni.fpc + C::TT - 1) / C::TT;
        kern<<<(unsigned)n_tiles, NT, C::smem_bytes, st>>>(wave, b->d_sample_off, b->d_frame_off, nullptr, (C::GT ? plan->d_window_half : plan->d_window),
                                                           plan->d_tw_half, plan->d_tw_full, plan->d_win_bq, plan->d_tw_kb, power, S,
                                                           reinterpret_cast<float2*>(cplx), uni);
    } else {
        // tile height: at most 16 frames, an even split of the longest clip
        int tt = C::TT;
        if (b->max_frames > 0) {
            const int64_t nt = (b->max_frames + tt - 1) / tt;
            tt = (int)((b->max_frames + nt - 1) / nt);
        }
        int rc = ensure_stft_tiles(b, tt);
        if (rc) return rc;
        if (b->n_stft_tiles == 0) return HPSS_OK;
        kern<<<b->n_stft_tiles, NT, C::smem_bytes, st>>>(wave, b->d_sample_off, b->d_frame_off, b->d_stft_tiles,
                                                         (C::GT ? plan->d_window_half : plan->d_window), plan->d_tw_half, plan->d_tw_full, plan->d_win_bq, plan->d_tw_kb, power, S,
                                                         reinterpret_cast<float2*>(cplx), uni);
    }
    HPSS_LAUNCHED("stft_fast_kernel");
    return HPSS_OK;
}

}  // namespace

bool stft_fast_split(int n_fft, int* na, int* nb) {
    switch (n_fft) {
        case 400: *na = 10; *nb = 20; return true;
        case 512: *na = 16; *nb = 16; return true;
        case 1024: *na = 16; *nb = 32; return true;
        case 2048: *na = 32; *nb = 32; return true;
        default: return false;
    }
}

// *handled = false (nothing launched) when (n_fft, hop) has no specialisation; the caller then runs the
// generic mixed-radix kernel.
int launch_stft_fast(hpss_ctx* ctx, hpss_batch* b, const float* wave, const FftPlan* plan, int hop, int power, float* S,
                     float* cplx, cudaStream_t st, bool* handled) {
    *handled = true;
    const int n = plan->n_fft;
    const bool use_r400 = knobs().k1_real != 0;
    if (n == 400 && hop == 160 && !power && !cplx && use_r400 && plan->d_win_r400) return launch_r400(ctx, b, wave, plan, S, st);
    if (n == 400 && hop == 160) return launch_cfg<400, 160, 10, 20, 10 * HPSS_K1_TT400>(ctx, b, wave, plan, power, S, cplx, st);
    if (n == 512 && hop == 160) return launch_cfg<512, 160, 16, 16, 256>(ctx, b, wave, plan, power, S, cplx, st);
    if (n == 512 && hop == 128) return launch_cfg<512, 128, 16, 16, 256>(ctx, b, wave, plan, power, S, cplx, st);
    if (n == 1024 && hop == 256) return launch_cfg<1024, 256, 16, 32, 256>(ctx, b, wave, plan, power, S, cplx, st);
    if (n == 2048 && hop == 512) return launch_cfg<2048, 512, 32, 32, 256>(ctx, b, wave, plan, power, S, cplx, st);
    *handled = false;
    return HPSS_OK;
}

}  // namespace hpss
