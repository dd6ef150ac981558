// K1: fused framing + window + shared-memory mixed-radix (2/3/4/5) Stockham FFT + |X|.
//
// Replaces librosa.core.stft(center=False) + np.abs (lib/preprocessing.py:381,387,407,
// 417,429,439 of the reference).  One CTA owns a tile of up to 16 consecutive frames of
// one clip: the overlapping sample segment is staged once in shared memory (each sample
// is read from HBM once per tile), every frame is transformed as a real-via-complex FFT
// of size n_fft/2 entirely in shared memory, and the magnitudes are written transposed
// so that consecutive lanes write consecutive frames of one frequency row.
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace hpss {

namespace {

// division by a launch-invariant small divisor: q = umulhi(x, ceil(2^32 / d)), exact while x*d < 2^32
struct FastDiv {
    uint32_t magic, d;
    __host__ __device__ FastDiv() : magic(0), d(1) {}
    __host__ explicit FastDiv(uint32_t dd) : magic(dd > 1 ? (uint32_t)((0x100000000ull + dd - 1) / dd) : 0u), d(dd) {}
    __device__ __forceinline__ uint32_t div(uint32_t x) const { return d == 1 ? x : __umulhi(x, magic); }
};

struct RadixList {
    int n_pass;
    int radix[kMaxRadixPasses];
    FastDiv div_m[kMaxRadixPasses];    // by n2 / radix[p]   (butterflies per frame)
    FastDiv div_ns[kMaxRadixPasses];   // by the product of the radices before pass p
    FastDiv div_nf;                    // by the number of frames in a full tile
    int m[kMaxRadixPasses];            // n2 / radix[p]
    int step[kMaxRadixPasses];         // n2 / (Ns * radix[p]): twiddle stride
    int fg[kMaxRadixPasses];           // frame groups running side by side: max(1, blockDim / m)
};

// sqrt.approx (MUFU): relative error ~1e-7, exact 0 for 0; magnitudes only need the 1e-4 feature tolerance
__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }

template <int R>
__device__ __forceinline__ void butterfly(float2 (&v)[R]);

template <>
__device__ __forceinline__ void butterfly<2>(float2 (&v)[2]) {
    const float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}
template <>
__device__ __forceinline__ void butterfly<3>(float2 (&v)[3]) {
    const float s = 0.86602540378443864676f;
    const float2 t1 = cadd(v[1], v[2]);
    const float2 d = csub(v[1], v[2]);
    const float2 t2 = make_float2(v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y);
    v[0] = cadd(v[0], t1);
    v[1] = make_float2(t2.x + s * d.y, t2.y - s * d.x);
    v[2] = make_float2(t2.x - s * d.y, t2.y + s * d.x);
}
template <>
__device__ __forceinline__ void butterfly<4>(float2 (&v)[4]) {
    const float2 t0 = cadd(v[0], v[2]);
    const float2 t1 = csub(v[0], v[2]);
    const float2 t2 = cadd(v[1], v[3]);
    const float2 t3 = mul_mi(csub(v[1], v[3]));
    v[0] = cadd(t0, t2);
    v[1] = cadd(t1, t3);
    v[2] = csub(t0, t2);
    v[3] = csub(t1, t3);
}
template <>
__device__ __forceinline__ void butterfly<5>(float2 (&v)[5]) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    const float2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]);
    const float2 b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
    const float2 p1 = make_float2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
    const float2 p2 = make_float2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
    const float2 q1 = make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
    const float2 q2 = make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
    v[0] = make_float2(v[0].x + a1.x + a2.x, v[0].y + a1.y + a2.y);
    const float2 m1 = mul_mi(q1), m2 = mul_mi(q2);
    v[1] = cadd(p1, m1);
    v[4] = csub(p1, m1);
    v[2] = cadd(p2, m2);
    v[3] = csub(p2, m2);
}

// One Stockham pass of radix R over `nf` frames.  FIRST: inputs come from the windowed sample segment
// (even samples -> real, odd -> imaginary part of the half-size sequence).
// A thread owns butterfly j for the frames fg, fg+FG, fg+2FG, ... : the index arithmetic, the R-1
// twiddles (or the R window pairs of the first pass) are set up once and reused for every frame.
template <int R, bool FIRST>
__device__ __forceinline__ void stockham_pass(int nf, int m, int step, int fg_full, int Ns, FastDiv dm, FastDiv dns,
                                              int hop, int zs,
                                              const float* __restrict__ s_samp,
                                              const float* __restrict__ s_win,
                                              const float2* __restrict__ s_tw,
                                              const float2* __restrict__ in, float2* __restrict__ out) {
    const int FG = min(fg_full, nf);   // frame groups running side by side
    const int nitems = m * FG;
    for (int idx = threadIdx.x; idx < nitems; idx += blockDim.x) {
        const int fg = (int)dm.div((uint32_t)idx);
        const int j = idx - fg * m;
        const int k = j - (int)dns.div((uint32_t)j) * Ns;
        const int d = (j - k) * R + k;
        float2 c[R];                       // FIRST: window pairs; else twiddles (c[0] unused)
#pragma unroll
        for (int q = 0; q < R; ++q) {
            if (FIRST) c[q] = *reinterpret_cast<const float2*>(s_win + 2 * (j + q * m));
            else if (q > 0) c[q] = s_tw[q * k * step];
        }
        for (int fr = fg; fr < nf; fr += FG) {
            float2 v[R];
            if (FIRST) {
                const float2* x = reinterpret_cast<const float2*>(s_samp + fr * hop) + j;   // hop is even
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    const float2 xv = x[q * m];
                    v[q] = make_float2(xv.x * c[q].x, xv.y * c[q].y);
                }
            } else {
                const float2* x = in + fr * zs + j;
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    v[q] = x[q * m];
                    if (q > 0) v[q] = cmul(v[q], c[q]);
                }
            }
            butterfly<R>(v);
            float2* y = out + fr * zs + d;
#pragma unroll
            for (int q = 0; q < R; ++q) y[q * Ns] = v[q];
        }
    }
}

template <bool FIRST>
__device__ __forceinline__ void run_pass(const RadixList& rl, int p, int nf, int Ns, int hop, int zs,
                                         const float* s_samp, const float* s_win, const float2* s_tw,
                                         const float2* in, float2* out) {
    const int m = rl.m[p], step = rl.step[p], fg = rl.fg[p];
    const FastDiv dm = rl.div_m[p], dns = rl.div_ns[p];
    switch (rl.radix[p]) {
        case 2: stockham_pass<2, FIRST>(nf, m, step, fg, Ns, dm, dns, hop, zs, s_samp, s_win, s_tw, in, out); break;
        case 3: stockham_pass<3, FIRST>(nf, m, step, fg, Ns, dm, dns, hop, zs, s_samp, s_win, s_tw, in, out); break;
        case 4: stockham_pass<4, FIRST>(nf, m, step, fg, Ns, dm, dns, hop, zs, s_samp, s_win, s_tw, in, out); break;
        default: stockham_pass<5, FIRST>(nf, m, step, fg, Ns, dm, dns, hop, zs, s_samp, s_win, s_tw, in, out); break;
    }
}

__global__ void __launch_bounds__(256)
stft_mag_kernel(const float* __restrict__ wave, const int64_t* __restrict__ sample_off,
                const int64_t* __restrict__ frame_off, const int2* __restrict__ tiles, int n_fft, int n2,
                int hop, int TT, RadixList rl, const float* __restrict__ window,
                const float2* __restrict__ tw_half, const float2* __restrict__ tw_full, int power,
                float* __restrict__ S, float2* __restrict__ cplx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int zs = n2 | 1;   // odd frame stride (in float2) -> conflict-free transposed reads
    const int seg_cap = (TT - 1) * hop + n_fft;
    float2* bufA = reinterpret_cast<float2*>(smem_raw);
    float2* bufB = bufA + (size_t)TT * zs;
    float2* s_twh = bufB + (size_t)TT * zs;
    float2* s_twf = s_twh + n2;
    float* s_win = reinterpret_cast<float*>(s_twf + (n2 + 1));
    float* s_samp = s_win + n_fft;

    const int2 tile = tiles[blockIdx.x];
    const int c = tile.x, t0 = tile.y;
    const int64_t fo = frame_off[c];
    const int T = (int)(frame_off[c + 1] - fo);
    const int nf = min(TT, T - t0);
    const int F = n2 + 1;

    // ---- stage tables and the sample segment
    for (int i = threadIdx.x; i < n_fft; i += blockDim.x) s_win[i] = window[i];
    for (int i = threadIdx.x; i < n2; i += blockDim.x) s_twh[i] = tw_half[i];
    for (int i = threadIdx.x; i <= n2; i += blockDim.x) s_twf[i] = tw_full[i];
    {
        const float* src = wave + sample_off[c] + (int64_t)t0 * hop;
        const int seg = (nf - 1) * hop + n_fft;
        // vector body when the global address is 16-byte aligned (pure offset otherwise)
        const int mis = (int)((reinterpret_cast<uintptr_t>(src) >> 2) & 3);
        const int head = min(seg, (4 - mis) & 3);
        for (int i = threadIdx.x; i < head; i += blockDim.x) s_samp[i] = __ldg(src + i);
        const int nvec = (seg - head) >> 2;
        const float4* src4 = reinterpret_cast<const float4*>(src + head);
        for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
            const float4 v = __ldg(src4 + i);
            float* d = s_samp + head + 4 * i;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
        for (int i = head + 4 * nvec + threadIdx.x; i < seg; i += blockDim.x) s_samp[i] = __ldg(src + i);
        (void)seg_cap;
    }
    __syncthreads();

    // ---- FFT passes (ping-pong between bufA and bufB)
    float2* in = bufB;
    float2* out = bufA;
    int Ns = 1;
    run_pass<true>(rl, 0, nf, Ns, hop, zs, s_samp, s_win, s_twh, in, out);
    Ns *= rl.radix[0];
    __syncthreads();
    for (int p = 1; p < rl.n_pass; ++p) {
        float2* t = in; in = out; out = t;
        run_pass<false>(rl, p, nf, Ns, hop, zs, s_samp, s_win, s_twh, in, out);
        Ns *= rl.radix[p];
        __syncthreads();
    }
    const float2* Z = out;

    // ---- real-FFT unpack + magnitude, written (f, t) with lanes along t.  Bins k and n2-k share their
    // loads: X[k] = E + W*O and X[n2-k] = conj(E - W*O) with E, O from Z[k] and conj(Z[n2-k]).
    const int64_t base = (int64_t)F * fo + t0;
    const int half = n2 / 2;
    const int fr = threadIdx.x & 15;            // tiles hold at most 16 frames
    if (fr < nf) {
        const float2* zrow = Z + fr * zs;
        float* Sg = S + base + fr;
        float2* Cg = cplx ? cplx + base + fr : nullptr;
        for (int k = threadIdx.x >> 4; k <= half; k += (int)(blockDim.x >> 4)) {
            const int k2 = n2 - k;
            const float2 zk = zrow[k];
            float2 zc = zrow[k == 0 ? 0 : k2];
            zc.y = -zc.y;
            const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
            const float2 dd = csub(zk, zc);
            const float2 o = make_float2(0.5f * dd.y, -0.5f * dd.x);
            const float2 wo = cmul(s_twf[k], o);
            const float2 xa = cadd(e, wo);                           // X[k]
            const float2 xb = make_float2(e.x - wo.x, wo.y - e.y);   // X[n2-k] = conj(E - W*O)
            const float pa = xa.x * xa.x + xa.y * xa.y;
            const float pb = xb.x * xb.x + xb.y * xb.y;
            const int64_t ga = (int64_t)k * T, gb = (int64_t)k2 * T;
            Sg[ga] = power ? pa : fast_sqrt(pa);
            if (Cg) Cg[ga] = xa;
            if (k2 != k) {
                Sg[gb] = power ? pb : fast_sqrt(pb);
                if (Cg) Cg[gb] = xb;
            }
        }
    }
}

}  // namespace

size_t stft_smem_bytes(int n_fft, int hop, int TT) {
    const int n2 = n_fft / 2;
    const int zs = n2 | 1;
    size_t b = 2 * (size_t)TT * zs * sizeof(float2);
    b += (size_t)n2 * sizeof(float2) + (size_t)(n2 + 1) * sizeof(float2);
    b += (size_t)n_fft * sizeof(float);
    b += ((size_t)(TT - 1) * hop + n_fft) * sizeof(float);
    return b;
}

// Tile height: at most 16 frames, at least 3 CTAs/SM worth of shared memory when possible,
// and an even split of the longest clip so the last tile is not mostly empty.
int choose_stft_tt(const hpss_ctx* ctx, const hpss_batch* b, int n_fft, int hop) {
    int tt = 16;
    const size_t budget3 = (size_t)(ctx->max_smem_optin + 1024) / 3 - 1024;
    while (tt > 1 && stft_smem_bytes(n_fft, hop, tt) > budget3) tt >>= 1;
    while (tt > 1 && stft_smem_bytes(n_fft, hop, tt) > (size_t)ctx->max_smem_optin) tt >>= 1;
    if (stft_smem_bytes(n_fft, hop, tt) > (size_t)ctx->max_smem_optin) return 0;
    const int64_t T = b->max_frames;
    if (T > 0) {
        const int64_t nt = (T + tt - 1) / tt;
        tt = (int)((T + nt - 1) / nt);
    }
    return tt;
}

int launch_stft(hpss_ctx* ctx, hpss_batch* b, const float* wave, const FftPlan* plan, int hop, int power,
                float* S, float* cplx, cudaStream_t st) {
    {   // specialised two-pass register FFT for the reference's (n_fft, hop) pairs
        const bool no_fast = knobs().no_fast_stft != 0;   // development knob
        bool handled = false;
        if (!no_fast) {
            const int rc = launch_stft_fast(ctx, b, wave, plan, hop, power, S, cplx, st, &handled);
            if (rc || handled) return rc;
        }
    }
    const int tt = choose_stft_tt(ctx, b, plan->n_fft, hop);
    if (tt <= 0) {
        set_error("n_fft=%d does not fit the shared-memory FFT (max %d bytes)", plan->n_fft,
                  ctx->max_smem_optin);
        return HPSS_ERR_UNSUPPORTED;
    }
    int rc = ensure_stft_tiles(b, tt);
    if (rc) return rc;
    if (b->n_stft_tiles == 0) return HPSS_OK;
    if (hop & 1) {
        set_error("hop_length=%d must be even (8-byte sample loads in the FFT's first pass)", hop);
        return HPSS_ERR_UNSUPPORTED;
    }
    RadixList rl;
    rl.n_pass = plan->n_pass;
    {
        int ns = 1;
        for (int i = 0; i < kMaxRadixPasses; ++i) {
            rl.radix[i] = plan->radix[i];
            rl.m[i] = rl.step[i] = rl.fg[i] = 1;
            if (i < plan->n_pass) {
                rl.m[i] = plan->n2 / plan->radix[i];
                rl.step[i] = plan->n2 / (ns * plan->radix[i]);
                rl.fg[i] = std::max(1, 256 / rl.m[i]);
                rl.div_m[i] = FastDiv((uint32_t)rl.m[i]);
                rl.div_ns[i] = FastDiv((uint32_t)ns);
                ns *= plan->radix[i];
            }
        }
        rl.div_nf = FastDiv((uint32_t)tt);
    }
    const size_t smem = stft_smem_bytes(plan->n_fft, hop, tt);
    HPSS_CUDA(cudaFuncSetAttribute(stft_mag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stft_mag_kernel<<<b->n_stft_tiles, 256, smem, st>>>(
        wave, b->d_sample_off, b->d_frame_off, b->d_stft_tiles, plan->n_fft, plan->n2, hop, tt, rl,
        plan->d_window, plan->d_tw_half, plan->d_tw_full, power, S, reinterpret_cast<float2*>(cplx));
    HPSS_LAUNCHED("stft_mag_kernel");
    return HPSS_OK;
}

}  // namespace hpss
