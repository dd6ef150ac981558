// N2: signal preparation on the GPU -- everything load_and_preprocess_signal does after the file is decoded,
// and mix_signals, for a whole batch of files per call.
//
// Reference: lib/preprocessing.py:330-350 (load_and_preprocess_signal), :114-132 (normalize_signal),
// :297-325 (mix_signals); lib/cython_impl/tools.pyx:42-134 (removeSilence, the Cython leaf the reference calls);
// librosa.feature.rms(y, frame_length, hop_length, center=True, pad_mode='reflect') as called at :337.
//
//   x (float32, or int16 PCM / 32768 as librosa.load returns it)
//     P1  per-clip sum / min / max                        -> mean1, peak1         (normalize_signal)
//     P2  frame energies sqrt(mean(y^2)), y = (x - mean1) / peak1, reflect-padded centred frames; per-clip max
//     P3  per clip: threshold alpha * max, 5-tap zero-padded median of the 0/1 markers, silent stretches
//         [k, l) longer than beta seconds; more than one of them -> the kept samples are packed to the front
//         of a buffer of ones that keeps the original length (tools.pyx:88-131)
//     P4  sum / min / max of that buffer                  -> mean2, peak2         (second normalize_signal)
//     P5  out = ((z - mean2) / peak2), repeated 2^r times while shorter than 0.1 s (:345-347)
//
// y is never stored: P2, P4 and P5 recompute it from x (two float32 operations), so the stage reads the waveform
// four times and writes it once: 20 B per sample for float32 input, 12 B for int16 PCM.  All sums are float64 and
// reduced in a fixed order (per-chunk partials, then one warp per clip), so results do not depend on scheduling.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace hpss {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kChunk = 8192;            // samples per CTA of the sample-level passes
constexpr int kPerThread = kChunk / kThreads;
constexpr int kBatch = 8;              // samples per thread in flight

struct PrepClip {        // host-built, one per clip
    int64_t in_off;      // first sample in the input buffer
    int64_t out_off;     // first sample in the output buffer
    int64_t fr_off;      // first entry of the per-frame arrays
    int32_t len;         // samples
    int32_t rep;         // output = rep copies back to back (the < 0.1 s doubling)
    int32_t n_frames;    // RMS frames (center=True)
    int32_t chunk0;      // first chunk of this clip
    int32_t n_chunks;
    int32_t pad_;
    int64_t fm_off;      // first entry of this clip in the caller's (compact) frame-marker array
};

struct RmsTile { int32_t clip, t0, nt, pad_; };     // frames [t0, t0 + nt) of one clip

struct PrepState {       // device-written, one per clip
    float mean1, peak1;
    uint32_t max_e;      // bits of the largest frame energy (non-negative floats order like their bits)
    int32_t n_sil;       // silent stretches longer than beta
    int32_t n_iv;        // intervals stored (== n_sil)
    int32_t apply;       // n_sil > 1: the compaction is applied
    int32_t removed;     // samples in the stored intervals
    int32_t n_kept;      // apply ? len - removed : len
    float mean2, peak2;
    float rcp1, rcp2;    // refined reciprocals of peak1 / peak2 (0 when the peak is outside [2^-60, 2^60]: exact path)
};

struct Partial { double sum; float mn, mx; int32_t bad; int32_t pad_; };

struct FrameArrays {     // all indexed by PrepClip::fr_off + i
    float* energy;
    uint8_t* marker;     // after the median
    int32_t* run_a;      // first frame of every zero run
    int32_t* run_b;      // first active frame after it (n_frames when the run reaches the end)
    int32_t* iv_k;       // qualifying intervals [k, l) in samples
    int32_t* iv_l;
    int32_t* iv_rb;      // samples removed before interval r
    int32_t* blk_first;  // per hop block h: number of intervals with l <= h * hop
};

template <typename T> __device__ __forceinline__ float load_sample(const T* p, int64_t i);
template <> __device__ __forceinline__ float load_sample<float>(const float* p, int64_t i) { return __ldg(p + i); }
template <> __device__ __forceinline__ float load_sample<int16_t>(const int16_t* p, int64_t i) {
    return (float)__ldg(p + i) * (1.0f / 32768.0f);      // exact: what librosa.load returns for 16-bit PCM
}

__device__ __forceinline__ float norm1(float x, float mean, float peak) {
    return __fdiv_rn(__fsub_rn(x, mean), peak);           // (Xin - mean) / max|.| in float32 (:130-131)
}

// refined reciprocal of a divisor that is used for a whole clip (the first half of the in-range sequence of an
// IEEE float32 division, maskmath.cuh::div_in_range); 0 when the divisor is outside [2^-60, 2^60]
__device__ __forceinline__ float refined_rcp(float p) {
    if (!(p >= 0x1p-60f && p <= 0x1p60f)) return 0.f;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p));
    const float e = __fmaf_rn(-p, r, 1.0f);
    return __fmaf_rn(r, e, r);
}
// (x - mean) / peak, correctly rounded like norm1: quotient, exact remainder, correction on the FMA pipe while
// the numerator is in range; the rare tiny / huge numerator (and a clip whose peak is out of range) divides exactly
__device__ __noinline__ float div_exact(float a, float p) { return __fdiv_rn(a, p); }   // cold: one copy in the code
__device__ __forceinline__ float norm_fast(float x, float mean, float peak, float r) {
    const float a = __fsub_rn(x, mean);
    // numerator (and divisor, see refined_rcp) in [2^-60, 2^60]: quotient, remainder and correction all stay in the
    // normal range.  One unsigned compare on the exponent field; zero takes the exact path too (it is rare and cheap).
    const uint32_t e = (__float_as_uint(a) >> 23) & 0xffu;
    if (r == 0.f || e - 67u > 120u) return div_exact(a, peak);
    const float q0 = __fmul_rn(a, r);
    const float rem = __fmaf_rn(-peak, q0, a);
    return __fmaf_rn(r, rem, q0);
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// numpy.pad(mode='reflect') index (whole-sample symmetric, the edge value is not repeated), any overshoot
__device__ __forceinline__ int reflect101(int i, int n) {
    if ((unsigned)i < (unsigned)n) return i;
    if (n == 1) return 0;
    const int p = 2 * (n - 1);
    i %= p;
    if (i < 0) i += p;
    return i < n ? i : p - i;
}

// position of sample s after the compaction; false when s lies in a removed interval.  `r` is the caller's
// running interval index (monotone in s).
struct Gate {
    const int32_t* k; const int32_t* l; const int32_t* rb;
    int n_iv, removed;
    int ck, cl, crb;          // interval r, cached: [ck, cl) and the samples removed before it (r == n_iv: none)
    __device__ __forceinline__ void fetch(int r) {
        if (r < n_iv) { ck = __ldg(k + r); cl = __ldg(l + r); crb = __ldg(rb + r); }
        else { ck = 0x7fffffff; cl = 0x7fffffff; crb = removed; }
    }
};
__device__ __forceinline__ bool gate_locate(Gate& g, int s, int& r, int& pos) {
    while (s >= g.cl) g.fetch(++r);
    if (s >= g.ck) return false;
    pos = s - g.crb;
    return true;
}

// ---- P1 / P4: per-chunk sum, min, max ------------------------------------------------------------------------
// MODE 0: x itself.  MODE 1: the buffer removeSilence returns (kept y, compaction applied or not; the tail of
// ones is added by the finalize kernel).
template <typename T, int MODE>
__global__ void __launch_bounds__(kThreads)
prep_reduce_kernel(const T* __restrict__ x, const PrepClip* __restrict__ clips, const int2* __restrict__ chunks,
                   const PrepState* __restrict__ state, FrameArrays fa, int hop, Partial* __restrict__ partial) {
    __shared__ double s_sum[kWarps];
    __shared__ float s_mn[kWarps], s_mx[kWarps];
    __shared__ int s_bad[kWarps];
    const int2 ch = __ldg(chunks + blockIdx.x);
    const PrepClip cl = clips[ch.x];
    const int s0 = ch.y, s1 = min(cl.len, ch.y + kChunk);
    const T* xp = x + cl.in_off;
    float mean1 = 0.f, peak1 = 1.f, rcp1 = 1.f;
    Gate g{nullptr, nullptr, nullptr, 0, 0, 0x7fffffff, 0x7fffffff, 0};
    int r = 0;
    if (MODE == 1) {
        const PrepState st = state[ch.x];
        mean1 = st.mean1; peak1 = st.peak1; rcp1 = st.rcp1;
        if (st.apply) {
            g = Gate{fa.iv_k + cl.fr_off, fa.iv_l + cl.fr_off, fa.iv_rb + cl.fr_off, st.n_iv, st.removed, 0, 0, 0};
            r = __ldg(fa.blk_first + cl.fr_off + s0 / hop);
            g.fetch(r);
        }
    }
    double sum = 0.0;
    float mn = INFINITY, mx = -INFINITY;
    int bad = 0;
    // batches of kBatch samples per thread: the loads of a batch are issued together, the arithmetic of a batch is
    // one basic block; the batch loop itself stays rolled (the unrolled kernel was four times the code and slower)
#pragma unroll 1
    for (int b0 = 0; b0 < kPerThread; b0 += kBatch) {
        float v[kBatch];
#pragma unroll
        for (int i = 0; i < kBatch; ++i) {
            const int s = s0 + threadIdx.x + (b0 + i) * kThreads;
            v[i] = s < s1 ? load_sample<T>(xp, s) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < kBatch; ++i) {
            const int s = s0 + threadIdx.x + (b0 + i) * kThreads;
            if (s >= s1) continue;
            float y = v[i];
            if (MODE == 0) {
                if (!isfinite(y)) { ++bad; continue; }
            } else {
                y = norm_fast(y, mean1, peak1, rcp1);
                int pos;
                if (g.n_iv && !gate_locate(g, s, r, pos)) continue;
            }
            sum += (double)y;
            mn = fminf(mn, y);
            mx = fmaxf(mx, y);
        }
    }
    sum = warp_sum_d(sum);
    mn = warp_min_f(mn);
    mx = warp_max_f(mx);
    bad = __reduce_add_sync(0xffffffffu, bad);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_sum[warp] = sum; s_mn[warp] = mn; s_mx[warp] = mx; s_bad[warp] = bad; }
    __syncthreads();
    if (threadIdx.x == 0) {
        Partial p{0.0, INFINITY, -INFINITY, 0, 0};
        for (int w = 0; w < kWarps; ++w) {
            p.sum += s_sum[w]; p.mn = fminf(p.mn, s_mn[w]); p.mx = fmaxf(p.mx, s_mx[w]); p.bad += s_bad[w];
        }
        partial[blockIdx.x] = p;
    }
}

// one warp per clip: fold the clip's chunk partials in a fixed order -> mean / peak of normalize_signal
template <int MODE>
__global__ void __launch_bounds__(kThreads)
prep_finalize_kernel(const PrepClip* __restrict__ clips, int n_clips, const Partial* __restrict__ partial,
                     PrepState* __restrict__ state, uint32_t* __restrict__ flags) {
    const int c = blockIdx.x * kWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= n_clips) return;
    const PrepClip cl = clips[c];
    double sum = 0.0;
    float mn = INFINITY, mx = -INFINITY;
    int bad = 0;
    for (int i = lane; i < cl.n_chunks; i += 32) {
        const Partial p = partial[cl.chunk0 + i];
        sum += p.sum; mn = fminf(mn, p.mn); mx = fmaxf(mx, p.mx); bad += p.bad;
    }
    sum = warp_sum_d(sum);
    mn = warp_min_f(mn);
    mx = warp_max_f(mx);
    bad = __reduce_add_sync(0xffffffffu, bad);
    if (lane != 0) return;
    PrepState st = state[c];
    if (MODE == 0) {
        if (bad) atomicOr(flags, 1u);                                     // HPSS_ERR_NONFINITE at the next hpss_ctx_check
        st.mean1 = (float)(sum / (double)cl.len);
        st.peak1 = fmaxf(fabsf(__fsub_rn(mx, st.mean1)), fabsf(__fsub_rn(mn, st.mean1)));
        st.rcp1 = refined_rcp(st.peak1);
        st.rcp2 = 0.f;
        st.max_e = 0u;
        st.n_sil = 0; st.n_iv = 0; st.apply = 0; st.removed = 0; st.n_kept = cl.len;
        st.mean2 = 0.f; st.peak2 = 1.f;
    } else {
        if (st.apply && st.removed > 0) {                                 // the tail of ones (tools.pyx:93, 128)
            sum += (double)st.removed;
            mn = fminf(mn, 1.f);
            mx = fmaxf(mx, 1.f);
        }
        st.mean2 = (float)(sum / (double)cl.len);
        st.peak2 = fmaxf(fabsf(__fsub_rn(mx, st.mean2)), fabsf(__fsub_rn(mn, st.mean2)));
        st.rcp2 = refined_rcp(st.peak2);
    }
    state[c] = st;
}

// ---- P2: librosa.feature.rms(center=True): one CTA per tile of consecutive frames of one clip.  The squares of the
// normalised samples of the tile's span ((nt - 1) * hop + win samples of the reflect-padded signal) are staged once in
// shared memory -- every sample is loaded, normalised and squared once, not once per overlapping frame -- and each
// warp then sums whole frames from there.
template <typename T>
__global__ void __launch_bounds__(kThreads)
prep_rms_kernel(const T* __restrict__ x, const PrepClip* __restrict__ clips, const RmsTile* __restrict__ tiles,
                PrepState* __restrict__ state, int win, int hop, float* __restrict__ energy) {
    extern __shared__ float s_sq[];
    __shared__ float s_max[kWarps];
    const RmsTile tl = tiles[blockIdx.x];
    const PrepClip cl = clips[tl.clip];
    const PrepState st = state[tl.clip];
    const T* xp = x + cl.in_off;
    const int span = (tl.nt - 1) * hop + win;
    const int first = tl.t0 * hop - win / 2;               // np.pad(y, frame_length // 2, mode='reflect')
    const bool interior = first >= 0 && first + span <= cl.len;
    // (batches of eight loads per thread were measured slower than this plain loop: 227 vs 184 us under ncu)
#pragma unroll 4
    for (int j = threadIdx.x; j < span; j += kThreads) {
        const int i = interior ? first + j : reflect101(first + j, cl.len);
        const float y = norm_fast(load_sample<T>(xp, i), st.mean1, st.peak1, st.rcp1);
        s_sq[j] = __fmul_rn(y, y);                         // np.abs(x)**2 rounds the square, then the sum
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float emax = 0.f;
    for (int t = warp; t < tl.nt; t += kWarps) {
        const float* p = s_sq + t * hop;
        float acc = 0.f;
        for (int j = lane; j < win; j += 32) acc = __fadd_rn(acc, p[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        const float e = __fsqrt_rn(__fdiv_rn(acc, (float)win));
        if (lane == 0) energy[cl.fr_off + tl.t0 + t] = e;
        emax = e > emax || e != e ? e : emax;              // a NaN energy poisons the maximum (and the threshold)
    }
    if (lane == 0) s_max[warp] = emax;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = 0.f;
        for (int w = 0; w < kWarps; ++w) m = (s_max[w] > m || s_max[w] != s_max[w]) ? s_max[w] : m;
        atomicMax(&state[tl.clip].max_e, __float_as_uint(m));    // e >= 0: non-negative floats order like their bits
    }
}

// ---- P3: markers, silent stretches, intervals (one CTA per clip) -----------------------------------------
constexpr int kGateThreads = 512;

// exclusive scan of one int per thread over the CTA; returns the exclusive prefix, `total` = sum over the CTA
__device__ __forceinline__ int block_exscan(int v, int* s_warp, int& total) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    __syncthreads();                       // s_warp may still be read from the previous call
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kGateThreads / 32; ++w) {
        const int x = s_warp[w];
        if (w < warp) base += x;
        tot += x;
    }
    total = tot;
    return base + inc - v;
}

__global__ void __launch_bounds__(kGateThreads)
prep_gate_kernel(const PrepClip* __restrict__ clips, PrepState* __restrict__ state, FrameArrays fa, int fs, int win,
                 int hop, double alpha, double beta, int32_t* __restrict__ frame_marker_out,
                 int32_t* __restrict__ n_sil_out) {
    __shared__ int s_warp[kGateThreads / 32];
    const int c = blockIdx.x;
    const PrepClip cl = clips[c];
    const int n = cl.n_frames, L = cl.len;
    const float* e = fa.energy + cl.fr_off;
    uint8_t* mk = fa.marker + cl.fr_off;
    int32_t* run_a = fa.run_a + cl.fr_off;
    int32_t* run_b = fa.run_b + cl.fr_off;
    int32_t* iv_k = fa.iv_k + cl.fr_off;
    int32_t* iv_l = fa.iv_l + cl.fr_off;
    int32_t* iv_rb = fa.iv_rb + cl.fr_off;
    int32_t* blk_first = fa.blk_first + cl.fr_off;
    // cdef float energyThresh = alpha * np.max(energy): a float64 product stored into a C float (tools.pyx:87)
    const float thr = (float)(alpha * (double)__uint_as_float(state[c].max_e));
    // frame_silMarker = medfilt(energy >= thr, 5) > 0.5: zero-padded window, at least three of five
    for (int f = threadIdx.x; f < n; f += kGateThreads) {
        int cnt = 0;
#pragma unroll
        for (int d = -2; d <= 2; ++d) {
            const int g = f + d;
            if (g >= 0 && g < n && e[g] >= thr) ++cnt;
        }
        const uint8_t m = cnt >= 3 ? 1 : 0;
        mk[f] = m;
        if (frame_marker_out) frame_marker_out[cl.fm_off + f] = m;
    }
    __syncthreads();
    // zero runs [a, b): the while loops of tools.pyx:103-113 visit exactly the maximal zero runs (the frame that ends
    // a run is active, so skipping it with i = j + 1 loses nothing)
    int n_a = 0, n_b = 0;
    for (int base = 0; base < n; base += kGateThreads) {
        const int f = base + threadIdx.x;
        const bool in = f < n;
        const int cur = in ? mk[f] : 1, prev = (in && f > 0) ? mk[f - 1] : 1;
        const int is_a = in && cur == 0 && (f == 0 || prev == 1);
        const int is_b = in && f > 0 && cur == 1 && prev == 0;
        int tot_a, tot_b;
        const int ex_a = block_exscan(is_a, s_warp, tot_a);
        const int ex_b = block_exscan(is_b, s_warp, tot_b);
        if (is_a) run_a[n_a + ex_a] = f;
        if (is_b) run_b[n_b + ex_b] = f;
        n_a += tot_a; n_b += tot_b;
    }
    if (threadIdx.x == 0 && n_b < n_a) run_b[n_b] = n;      // the last run reaches the end of the clip
    __syncthreads();
    // k = max(hop*(i-1)+win, 1), l = min(hop*(j-1)+win, nSamples); kept when (l-k)/fs > beta (tools.pyx:114-124);
    // j stops at the last frame when the run reaches the end (the `if j == nFrames-1: break` of :110-112)
    int n_q = 0, removed = 0;
    for (int base = 0; base < n_a; base += kGateThreads) {
        const int r = base + threadIdx.x;
        int q = 0, k = 0, l = 0;
        if (r < n_a) {
            const int a = run_a[r], b = run_b[r];
            const int j = b < n ? b : n - 1;
            k = max(hop * (a - 1) + win, 1);
            l = min(hop * (j - 1) + win, L);
            q = ((double)(l - k) / (double)fs > beta) ? 1 : 0;
        }
        int tot_q, tot_len;
        const int ex_q = block_exscan(q, s_warp, tot_q);
        const int ex_len = block_exscan(q ? l - k : 0, s_warp, tot_len);
        if (q) { iv_k[n_q + ex_q] = k; iv_l[n_q + ex_q] = l; iv_rb[n_q + ex_q] = removed + ex_len; }
        n_q += tot_q; removed += tot_len;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        PrepState st = state[c];
        st.n_sil = n_q; st.n_iv = n_q; st.removed = removed;
        st.apply = n_q > 1 ? 1 : 0;                          // `if nSil > 1` (tools.pyx:127)
        st.n_kept = st.apply ? L - removed : L;
        state[c] = st;
        if (n_sil_out) n_sil_out[c] = n_q;
    }
    // per hop block: intervals that end at or before its first sample (entry point of the per-sample walk)
    const int n_blk = (L + hop - 1) / hop;
    for (int h = threadIdx.x; h < n_blk; h += kGateThreads) {
        const int s = h * hop;
        int lo = 0, hi = n_q;                                // first r with iv_l[r] > s
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (iv_l[mid] <= s) lo = mid + 1; else hi = mid;
        }
        blk_first[h] = lo;
    }
}

// ---- P5: write the prepared signal ----------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
prep_write_kernel(const T* __restrict__ x, const PrepClip* __restrict__ clips, const int2* __restrict__ chunks,
                  const PrepState* __restrict__ state, FrameArrays fa, int hop, float* __restrict__ out,
                  uint8_t* __restrict__ sample_marker) {
    const int2 ch = __ldg(chunks + blockIdx.x);
    const PrepClip cl = clips[ch.x];
    const PrepState st = state[ch.x];
    const int s0 = ch.y, s1 = min(cl.len, ch.y + kChunk);
    const T* xp = x + cl.in_off;
    float* op = out + cl.out_off;
    Gate g{fa.iv_k + cl.fr_off, fa.iv_l + cl.fr_off, fa.iv_rb + cl.fr_off, st.n_iv, st.removed, 0, 0, 0};
    int r = st.n_iv ? __ldg(fa.blk_first + cl.fr_off + s0 / hop) : 0;
    g.fetch(r);
    const float tail = norm1(1.0f, st.mean2, st.peak2);
#pragma unroll 1
    for (int b0 = 0; b0 < kPerThread; b0 += kBatch) {
        float v[kBatch];
#pragma unroll
        for (int i = 0; i < kBatch; ++i) {
            const int s = s0 + threadIdx.x + (b0 + i) * kThreads;
            v[i] = s < s1 ? load_sample<T>(xp, s) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < kBatch; ++i) {
            const int s = s0 + threadIdx.x + (b0 + i) * kThreads;
            if (s >= s1) continue;
            const float z = norm_fast(norm_fast(v[i], st.mean1, st.peak1, st.rcp1), st.mean2, st.peak2, st.rcp2);
            int pos = s;
            bool kept = true;
            if (st.n_iv) kept = gate_locate(g, s, r, pos);
            if (sample_marker) sample_marker[cl.in_off + s] = kept ? 1 : 0;
            if (!st.apply) { pos = s; }
            if (kept || !st.apply) {
                op[pos] = z;
                if (cl.rep > 1) {                       // clips shorter than 0.1 s: the doubling (rare)
#pragma unroll 1
                    for (int q = 1; q < cl.rep; ++q) op[(int64_t)q * cl.len + pos] = z;
                }
            }
            if (st.apply && s >= st.n_kept) {
                op[s] = tail;
                if (cl.rep > 1) {
#pragma unroll 1
                    for (int q = 1; q < cl.rep; ++q) op[(int64_t)q * cl.len + s] = tail;
                }
            }
        }
    }
}

// ---- mix_signals (lib/preprocessing.py:297-325) -----------------------------------------------------------------
// Pair p: speech sp[0, n) and music mu looped to n samples.  Energies, gains and the final normalisation are
// evaluated in float64 and rounded to float32 once (numpy >= 2 evaluates the reference's expression in float64
// because the gains are float64 scalars; numpy 1 kept float32 -- the two differ by float32 rounding).
struct MixPair {
    int64_t sp_off, mu_off, out_off;
    int32_t n, n_mu;
    int32_t chunk0, n_chunks;
    double target_db;
};
struct MixState { double g_sp, g_mu, mean, peak; };
struct MixPartial { double a, b; double mn, mx; };

template <int MODE>     // 0: sum sp^2, sum mu^2;  1: sum / min / max of the mix
__global__ void __launch_bounds__(kThreads)
mix_reduce_kernel(const float* __restrict__ sp, const float* __restrict__ mu, const MixPair* __restrict__ pairs,
                  const int2* __restrict__ chunks, const MixState* __restrict__ state, MixPartial* __restrict__ partial) {
    __shared__ double s_a[kWarps], s_b[kWarps], s_mn[kWarps], s_mx[kWarps];
    const int2 ch = __ldg(chunks + blockIdx.x);
    const MixPair pr = pairs[ch.x];
    const int s0 = ch.y, s1 = min(pr.n, ch.y + kChunk);
    const float* a = sp + pr.sp_off;
    const float* b = mu + pr.mu_off;
    double g_sp = 0, g_mu = 0;
    if (MODE == 1) { g_sp = state[ch.x].g_sp; g_mu = state[ch.x].g_mu; }
    double acc_a = 0, acc_b = 0, mn = INFINITY, mx = -INFINITY;
    for (int s = s0 + threadIdx.x; s < s1; s += kThreads) {
        const double x = (double)__ldg(a + s), y = (double)__ldg(b + s % pr.n_mu);
        if (MODE == 0) { acc_a += x * x; acc_b += y * y; }
        else { const double m = g_sp * x + g_mu * y; acc_a += m; mn = fmin(mn, m); mx = fmax(mx, m); }
    }
    acc_a = warp_sum_d(acc_a); acc_b = warp_sum_d(acc_b);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_a[warp] = acc_a; s_b[warp] = acc_b; s_mn[warp] = mn; s_mx[warp] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        MixPartial p{0, 0, INFINITY, -INFINITY};
        for (int w = 0; w < kWarps; ++w) { p.a += s_a[w]; p.b += s_b[w]; p.mn = fmin(p.mn, s_mn[w]); p.mx = fmax(p.mx, s_mx[w]); }
        partial[blockIdx.x] = p;
    }
}

template <int MODE>
__global__ void __launch_bounds__(kThreads)
mix_finalize_kernel(const MixPair* __restrict__ pairs, int n_pairs, const MixPartial* __restrict__ partial,
                    MixState* __restrict__ state) {
    const int p = blockIdx.x * kWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (p >= n_pairs) return;
    const MixPair pr = pairs[p];
    double a = 0, b = 0, mn = INFINITY, mx = -INFINITY;
    for (int i = lane; i < pr.n_chunks; i += 32) {
        const MixPartial q = partial[pr.chunk0 + i];
        a += q.a; b += q.b; mn = fmin(mn, q.mn); mx = fmax(mx, q.mx);
    }
    a = warp_sum_d(a); b = warp_sum_d(b);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane != 0) return;
    MixState st = state[p];
    if (MODE == 0) {
        const double e_sp = a / (double)pr.n, e_mu = b / (double)pr.n;              // :311-312
        const double req = e_sp / pow(10.0, pr.target_db / 10.0);                    // :314
        double g_mu = sqrt(req / e_mu);                                              // :315
        const double tot = g_mu + 1.0;                                               // :317-320
        st.g_mu = g_mu / tot;
        st.g_sp = 1.0 / tot;
        st.mean = 0; st.peak = 1;
    } else {
        st.mean = a / (double)pr.n;
        st.peak = fmax(fabs(mx - st.mean), fabs(mn - st.mean));
    }
    state[p] = st;
}

__global__ void __launch_bounds__(kThreads)
mix_write_kernel(const float* __restrict__ sp, const float* __restrict__ mu, const MixPair* __restrict__ pairs,
                 const int2* __restrict__ chunks, const MixState* __restrict__ state, float* __restrict__ out) {
    const int2 ch = __ldg(chunks + blockIdx.x);
    const MixPair pr = pairs[ch.x];
    const MixState st = state[ch.x];
    const int s0 = ch.y, s1 = min(pr.n, ch.y + kChunk);
    const float* a = sp + pr.sp_off;
    const float* b = mu + pr.mu_off;
    float* o = out + pr.out_off;
    for (int s = s0 + threadIdx.x; s < s1; s += kThreads) {
        const double m = st.g_sp * (double)__ldg(a + s) + st.g_mu * (double)__ldg(b + s % pr.n_mu);
        o[s] = (float)((m - st.mean) / st.peak);
    }
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// validation passes: set bit `bit` of the context status word when a value is non-finite (MODE 0,
// librosa.util.valid_audio) or negative (MODE 1, librosa.util.softmask's input check)
template <int MODE>
__global__ void __launch_bounds__(kThreads)
check_kernel(const float* __restrict__ x, int64_t n, uint32_t* __restrict__ flags, uint32_t bit) {
    bool bad = false;
    const int64_t n4 = n / 4;
    const bool aligned = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    auto test = [](float v) { return MODE == 0 ? !isfinite(v) : (v < 0.f); };
    if (aligned) {
        const float4* x4 = reinterpret_cast<const float4*>(x);
        for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kThreads) {
            const float4 v = __ldg(x4 + i);
            bad |= test(v.x) | test(v.y) | test(v.z) | test(v.w);
        }
        for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
            bad |= test(__ldg(x + i));
    } else {
        for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
            bad |= test(__ldg(x + i));
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flags, bit);
}

}  // namespace

int launch_check(hpss_ctx* ctx, const float* x, int64_t n, int mode, cudaStream_t st) {
    if (n <= 0) return HPSS_OK;
    int64_t grid = (n / 4 + kThreads - 1) / kThreads + 1;
    if (grid > (int64_t)ctx->sm_count * 16) grid = (int64_t)ctx->sm_count * 16;
    if (mode == 0) check_kernel<0><<<(unsigned)grid, kThreads, 0, st>>>(x, n, ctx->d_flags, 1u);
    else check_kernel<1><<<(unsigned)grid, kThreads, 0, st>>>(x, n, ctx->d_flags, 2u);
    HPSS_LAUNCHED("check_kernel");
    return HPSS_OK;
}

int64_t prep_out_length(int64_t n, int fs) {
    if (n <= 0) return 0;
    int64_t len = n;
    while ((double)len / (double)fs < 0.1) len *= 2;          // lib/preprocessing.py:345-347
    return len;
}

int64_t prep_num_frames(int64_t n, int win, int hop) {
    if (n <= 0 || win < 1 || hop < 1) return 0;
    const int64_t padded = n + 2 * (int64_t)(win / 2);
    return padded >= win ? 1 + (padded - win) / hop : 0;
}

static int ensure_prep_scratch(hpss_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->prep_ws_bytes) return HPSS_OK;
    if (ctx->prep_ws) {
        HPSS_CUDA(cudaDeviceSynchronize());
        HPSS_CUDA(cudaFree(ctx->prep_ws));
        ctx->prep_ws = nullptr; ctx->prep_ws_bytes = 0;
    }
    const size_t want = bytes + bytes / 4;
    cudaError_t e = cudaMalloc(&ctx->prep_ws, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("out of device memory: signal-preparation scratch of %zu bytes", want);
        return HPSS_ERR_NOMEM;
    }
    ctx->prep_ws_bytes = want;
    return HPSS_OK;
}

// Carves `n` typed arrays out of a byte cursor.
template <typename T> static T* carve(char*& p, size_t count) {
    T* r = (T*)p;
    p += align256(sizeof(T) * std::max<size_t>(count, 1));
    return r;
}

namespace {

struct PrepHost {                  // host-side layout of one batch of files
    std::vector<PrepClip> clips;
    std::vector<int2> chunks;
    std::vector<RmsTile> tiles;
    size_t n_fr = 0;
    int rms_tt = 1;                // frames per RMS tile (upper bound)
};

// frames per RMS tile: at most 32, and a span of at most ~12 K floats of shared memory
int rms_tile_frames(int win, int hop) {
    int tt = win < 12000 ? (12000 - win) / hop + 1 : 1;
    return std::max(1, std::min(32, tt));
}

int prep_layout(const int64_t* clip_len, int n_clips, int fs, int win, int hop, PrepHost& h) {
    h.clips.resize(n_clips);
    h.chunks.clear();
    h.tiles.clear();
    h.rms_tt = rms_tile_frames(win, hop);
    int64_t in_off = 0, out_off = 0, fr_off = 0, fm_off = 0;
    for (int c = 0; c < n_clips; ++c) {
        const int64_t L = clip_len[c];
        if (L < 2 || L > 0x7fffffffLL / 2) {
            set_error("prep_signals: clip %d has %lld samples (need 2 .. 2^30)", c, (long long)L);
            return HPSS_ERR_INVALID;
        }
        PrepClip& cl = h.clips[c];
        cl.in_off = in_off; cl.out_off = out_off; cl.fr_off = fr_off;
        cl.len = (int32_t)L;
        const int64_t ol = prep_out_length(L, fs);
        cl.rep = (int32_t)(ol / L);
        cl.n_frames = (int32_t)prep_num_frames(L, win, hop);
        cl.chunk0 = (int32_t)h.chunks.size();
        cl.n_chunks = (int32_t)((L + kChunk - 1) / kChunk);
        cl.pad_ = 0;
        cl.fm_off = fm_off;
        fm_off += cl.n_frames;
        for (int64_t s = 0; s < L; s += kChunk) h.chunks.push_back(make_int2(c, (int)s));
        {   // RMS tiles: an even split of the clip's frames into tiles of at most rms_tt frames
            const int nt = (cl.n_frames + h.rms_tt - 1) / h.rms_tt;
            const int tt = nt > 0 ? (cl.n_frames + nt - 1) / nt : 0;
            for (int t0 = 0; t0 < cl.n_frames; t0 += tt) h.tiles.push_back(RmsTile{c, t0, std::min(tt, cl.n_frames - t0), 0});
        }
        in_off += L; out_off += ol;
        // the per-frame arrays also serve as per-hop-block table: a clip owns max(n_frames, ceil(L/hop)) + 1 entries,
        // of which [fr_off, fr_off + n_frames) are real frames
        fr_off += std::max<int64_t>(cl.n_frames, (L + hop - 1) / hop) + 1;
    }
    h.n_fr = (size_t)fr_off;
    return HPSS_OK;
}

size_t desc_bytes(int n_clips, size_t n_chunks, size_t n_tiles) {
    return align256(sizeof(PrepClip) * std::max(n_clips, 1)) + align256(sizeof(int2) * std::max<size_t>(n_chunks, 1)) +
           align256(sizeof(RmsTile) * std::max<size_t>(n_tiles, 1));
}
size_t work_bytes(int n_clips, size_t n_chunks, size_t n_fr) {
    return align256(sizeof(PrepState) * std::max(n_clips, 1)) + align256(sizeof(Partial) * std::max<size_t>(n_chunks, 1)) +
           align256(sizeof(float) * std::max<size_t>(n_fr, 1)) + align256(std::max<size_t>(n_fr, 1)) +
           6 * align256(sizeof(int32_t) * std::max<size_t>(n_fr, 1));
}

// the seven launches; `desc` holds clips | chunks | fr_off, `work` the per-call scratch
int prep_run(hpss_ctx* ctx, char* desc, char* work, int n_clips, size_t n_chunks, size_t n_tiles, size_t n_fr,
             const void* pcm, int pcm_format, int fs, int win, int hop, double alpha, double beta, float* out,
             int32_t* frame_marker, uint8_t* sample_marker, int32_t* n_sil, cudaStream_t st) {
    PrepClip* d_clips = carve<PrepClip>(desc, n_clips);
    int2* d_chunks = carve<int2>(desc, n_chunks);
    RmsTile* d_tiles = carve<RmsTile>(desc, n_tiles);
    PrepState* d_state = carve<PrepState>(work, n_clips);
    Partial* d_partial = carve<Partial>(work, n_chunks);
    FrameArrays fa;
    fa.energy = carve<float>(work, n_fr);
    fa.marker = carve<uint8_t>(work, n_fr);
    fa.run_a = carve<int32_t>(work, n_fr);
    fa.run_b = carve<int32_t>(work, n_fr);
    fa.iv_k = carve<int32_t>(work, n_fr);
    fa.iv_l = carve<int32_t>(work, n_fr);
    fa.iv_rb = carve<int32_t>(work, n_fr);
    fa.blk_first = carve<int32_t>(work, n_fr);
    const unsigned g_chunks = (unsigned)n_chunks, g_clips = (unsigned)((n_clips + kWarps - 1) / kWarps);
    const size_t rms_smem = sizeof(float) * ((size_t)(rms_tile_frames(win, hop) - 1) * hop + win);
    const bool s16 = pcm_format == HPSS_PCM_S16;
    if (rms_smem > 48 * 1024) {
        if (s16) HPSS_CUDA(cudaFuncSetAttribute(prep_rms_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rms_smem));
        else HPSS_CUDA(cudaFuncSetAttribute(prep_rms_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rms_smem));
    }
    if (s16) prep_reduce_kernel<int16_t, 0><<<g_chunks, kThreads, 0, st>>>((const int16_t*)pcm, d_clips, d_chunks, d_state, fa, hop, d_partial);
    else prep_reduce_kernel<float, 0><<<g_chunks, kThreads, 0, st>>>((const float*)pcm, d_clips, d_chunks, d_state, fa, hop, d_partial);
    HPSS_LAUNCHED("prep_reduce_kernel");
    prep_finalize_kernel<0><<<g_clips, kThreads, 0, st>>>(d_clips, n_clips, d_partial, d_state, ctx->d_flags);
    HPSS_LAUNCHED("prep_finalize_kernel");
    if (s16) prep_rms_kernel<int16_t><<<(unsigned)n_tiles, kThreads, rms_smem, st>>>((const int16_t*)pcm, d_clips, d_tiles, d_state, win, hop, fa.energy);
    else prep_rms_kernel<float><<<(unsigned)n_tiles, kThreads, rms_smem, st>>>((const float*)pcm, d_clips, d_tiles, d_state, win, hop, fa.energy);
    HPSS_LAUNCHED("prep_rms_kernel");
    prep_gate_kernel<<<(unsigned)n_clips, kGateThreads, 0, st>>>(d_clips, d_state, fa, fs, win, hop, alpha, beta, frame_marker, n_sil);
    HPSS_LAUNCHED("prep_gate_kernel");
    if (s16) prep_reduce_kernel<int16_t, 1><<<g_chunks, kThreads, 0, st>>>((const int16_t*)pcm, d_clips, d_chunks, d_state, fa, hop, d_partial);
    else prep_reduce_kernel<float, 1><<<g_chunks, kThreads, 0, st>>>((const float*)pcm, d_clips, d_chunks, d_state, fa, hop, d_partial);
    HPSS_LAUNCHED("prep_reduce_kernel");
    prep_finalize_kernel<1><<<g_clips, kThreads, 0, st>>>(d_clips, n_clips, d_partial, d_state, ctx->d_flags);
    HPSS_LAUNCHED("prep_finalize_kernel");
    if (s16) prep_write_kernel<int16_t><<<g_chunks, kThreads, 0, st>>>((const int16_t*)pcm, d_clips, d_chunks, d_state, fa, hop, out, sample_marker);
    else prep_write_kernel<float><<<g_chunks, kThreads, 0, st>>>((const float*)pcm, d_clips, d_chunks, d_state, fa, hop, out, sample_marker);
    HPSS_LAUNCHED("prep_write_kernel");
    return HPSS_OK;
}

int upload_desc(const PrepHost& h, char* desc, cudaStream_t st, bool sync) {
    const int n_clips = (int)h.clips.size();
    PrepClip* d_clips = carve<PrepClip>(desc, n_clips);
    int2* d_chunks = carve<int2>(desc, h.chunks.size());
    RmsTile* d_tiles = carve<RmsTile>(desc, h.tiles.size());
    if (sync) {
        HPSS_CUDA(cudaMemcpy(d_clips, h.clips.data(), sizeof(PrepClip) * n_clips, cudaMemcpyHostToDevice));
        HPSS_CUDA(cudaMemcpy(d_chunks, h.chunks.data(), sizeof(int2) * h.chunks.size(), cudaMemcpyHostToDevice));
        HPSS_CUDA(cudaMemcpy(d_tiles, h.tiles.data(), sizeof(RmsTile) * h.tiles.size(), cudaMemcpyHostToDevice));
    } else {
        // pageable sources: the copies have consumed the host vectors when the calls return
        HPSS_CUDA(cudaMemcpyAsync(d_clips, h.clips.data(), sizeof(PrepClip) * n_clips, cudaMemcpyHostToDevice, st));
        HPSS_CUDA(cudaMemcpyAsync(d_chunks, h.chunks.data(), sizeof(int2) * h.chunks.size(), cudaMemcpyHostToDevice, st));
        HPSS_CUDA(cudaMemcpyAsync(d_tiles, h.tiles.data(), sizeof(RmsTile) * h.tiles.size(), cudaMemcpyHostToDevice, st));
    }
    return HPSS_OK;
}

}  // namespace

// one-shot: descriptors and scratch both live in the context's preparation scratch
int launch_prep(hpss_ctx* ctx, const void* pcm, int pcm_format, const int64_t* clip_len, int n_clips, int fs, int win,
                int hop, double alpha, double beta, float* out, int32_t* frame_marker, uint8_t* sample_marker,
                int32_t* n_sil, cudaStream_t st) {
    if (n_clips == 0) return HPSS_OK;
    PrepHost h;
    int rc = prep_layout(clip_len, n_clips, fs, win, hop, h);
    if (rc) return rc;
    const size_t db = desc_bytes(n_clips, h.chunks.size(), h.tiles.size());
    rc = ensure_prep_scratch(ctx, db + work_bytes(n_clips, h.chunks.size(), h.n_fr) + 4096);
    if (rc) return rc;
    char* desc = (char*)ctx->prep_ws;
    rc = upload_desc(h, desc, st, false);
    if (rc) return rc;
    return prep_run(ctx, desc, desc + db, n_clips, h.chunks.size(), h.tiles.size(), h.n_fr, pcm, pcm_format, fs, win, hop,
                    alpha, beta, out, frame_marker, sample_marker, n_sil, st);
}

// prebuilt descriptors for repeated use (the host pipeline): nothing is uploaded on the launch path
int prep_plan_build(hpss_ctx* ctx, const int64_t* clip_len, int n_clips, int fs, int win, int hop, PrepPlan** out) {
    *out = nullptr;
    PrepHost h;
    int rc = prep_layout(clip_len, n_clips, fs, win, hop, h);
    if (rc) return rc;
    PrepPlan* pp = new PrepPlan();
    pp->n_clips = n_clips; pp->n_chunks = h.chunks.size(); pp->n_tiles = h.tiles.size(); pp->n_fr = h.n_fr;
    pp->work_bytes = work_bytes(n_clips, h.chunks.size(), h.n_fr) + 4096;
    cudaError_t e = cudaMalloc(&pp->d_desc, desc_bytes(n_clips, h.chunks.size(), h.tiles.size()));
    if (e != cudaSuccess) { cudaGetLastError(); delete pp; set_error("out of device memory: preparation descriptors"); return HPSS_ERR_NOMEM; }
    rc = upload_desc(h, (char*)pp->d_desc, nullptr, true);
    if (rc) { cudaFree(pp->d_desc); delete pp; return rc; }
    rc = ensure_prep_scratch(ctx, pp->work_bytes);
    if (rc) { cudaFree(pp->d_desc); delete pp; return rc; }
    *out = pp;
    return HPSS_OK;
}

void prep_plan_free(PrepPlan* pp) {
    if (!pp) return;
    if (pp->d_desc) cudaFree(pp->d_desc);
    delete pp;
}

int launch_prep_plan(hpss_ctx* ctx, const PrepPlan* pp, const void* pcm, int pcm_format, int fs, int win, int hop,
                     double alpha, double beta, float* out, cudaStream_t st) {
    if (pp->n_clips == 0) return HPSS_OK;
    int rc = ensure_prep_scratch(ctx, pp->work_bytes);     // no-op unless another caller shrank nothing: grow-only
    if (rc) return rc;
    return prep_run(ctx, (char*)pp->d_desc, (char*)ctx->prep_ws, pp->n_clips, pp->n_chunks, pp->n_tiles, pp->n_fr, pcm,
                    pcm_format, fs, win, hop, alpha, beta, out, nullptr, nullptr, nullptr, st);
}

int launch_mix(hpss_ctx* ctx, const float* sp, const int64_t* sp_len, const float* mu, const int64_t* mu_len,
               const double* target_db, int n_pairs, float* out, cudaStream_t st) {
    if (n_pairs == 0) return HPSS_OK;
    std::vector<MixPair> pairs(n_pairs);
    std::vector<int2> chunks;
    int64_t so = 0, mo = 0;
    for (int i = 0; i < n_pairs; ++i) {
        if (sp_len[i] < 1 || mu_len[i] < 1 || sp_len[i] > 0x7fffffffLL || mu_len[i] > 0x7fffffffLL) {
            set_error("mix_signals: pair %d has %lld / %lld samples", i, (long long)sp_len[i], (long long)mu_len[i]);
            return HPSS_ERR_INVALID;
        }
        MixPair& pr = pairs[i];
        pr.sp_off = so; pr.mu_off = mo; pr.out_off = so;
        pr.n = (int32_t)sp_len[i]; pr.n_mu = (int32_t)mu_len[i];
        pr.chunk0 = (int32_t)chunks.size();
        pr.n_chunks = (int32_t)((sp_len[i] + kChunk - 1) / kChunk);
        pr.target_db = target_db[i];
        for (int64_t s = 0; s < sp_len[i]; s += kChunk) chunks.push_back(make_int2(i, (int)s));
        so += sp_len[i]; mo += mu_len[i];
    }
    const size_t n_chunks = chunks.size();
    const size_t bytes = align256(sizeof(MixPair) * n_pairs) + align256(sizeof(int2) * n_chunks) +
                         align256(sizeof(MixState) * n_pairs) + align256(sizeof(MixPartial) * n_chunks);
    int rc = ensure_prep_scratch(ctx, bytes);
    if (rc) return rc;
    char* p = (char*)ctx->prep_ws;
    MixPair* d_pairs = carve<MixPair>(p, n_pairs);
    int2* d_chunks = carve<int2>(p, n_chunks);
    MixState* d_state = carve<MixState>(p, n_pairs);
    MixPartial* d_partial = carve<MixPartial>(p, n_chunks);
    HPSS_CUDA(cudaMemcpyAsync(d_pairs, pairs.data(), sizeof(MixPair) * n_pairs, cudaMemcpyHostToDevice, st));
    HPSS_CUDA(cudaMemcpyAsync(d_chunks, chunks.data(), sizeof(int2) * n_chunks, cudaMemcpyHostToDevice, st));
    const unsigned g_chunks = (unsigned)n_chunks, g_pairs = (unsigned)((n_pairs + kWarps - 1) / kWarps);
    mix_reduce_kernel<0><<<g_chunks, kThreads, 0, st>>>(sp, mu, d_pairs, d_chunks, d_state, d_partial);
    HPSS_LAUNCHED("mix_reduce_kernel");
    mix_finalize_kernel<0><<<g_pairs, kThreads, 0, st>>>(d_pairs, n_pairs, d_partial, d_state);
    HPSS_LAUNCHED("mix_finalize_kernel");
    mix_reduce_kernel<1><<<g_chunks, kThreads, 0, st>>>(sp, mu, d_pairs, d_chunks, d_state, d_partial);
    HPSS_LAUNCHED("mix_reduce_kernel");
    mix_finalize_kernel<1><<<g_pairs, kThreads, 0, st>>>(d_pairs, n_pairs, d_partial, d_state);
    HPSS_LAUNCHED("mix_finalize_kernel");
    mix_write_kernel<<<g_chunks, kThreads, 0, st>>>(sp, mu, d_pairs, d_chunks, d_state, out);
    HPSS_LAUNCHED("mix_write_kernel");
    return HPSS_OK;
}

}  // namespace hpss
