// K5 + N1: feature moments for get_data_stats, scale_data, per-clip row standardisation
// (sklearn StandardScaler as used by get_feature_patches) and the patch gather of
// lib/cython_impl/tools.pyx:extract_patches.
//
// Reference: lib/preprocessing.py:461-586 (two-pass corpus statistics), :590-614 and
// tools.pyx:138-166 (scale_data), :137-292 + tools.pyx:21-38 (patches).
#include <stdlib.h>

#include <algorithm>

#include "common.cuh"

namespace hpss {

namespace {

#ifndef HPSS_MOM_MINB
#define HPSS_MOM_MINB 3
#endif
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Raw moments.  grid = (frame chunks, groups of 8 row pairs): every warp owns TWO feature rows (d and
// d + ceil(D/2): for the HPSS features the same band of the harmonic and the percussive stream) for its whole
// life and walks the clips of its frame chunk in blocks of 128 frames (lanes along time, four coalesced
// 128-byte reads per row and block, eight loads in flight per warp; the per-clip set-up -- offsets, class,
// thresholds -- is shared by the two rows).  Within a block every lane adds at most four values in float32;
// the block's partial sum and sum of squares are then folded into per-lane float64 accumulators (one sum per
// class + the sum of squares, per row) that stay in registers and are reduced across the warp once, at the
// end: 1 + n_classes float64 atomics per row and warp.
// A non-finite value makes the block's float32 sum of squares non-finite; the block is then redone value by
// value (non-finite values are counted and contribute 0).
// CLIP: fused K3b -- x = max(x, max_clip_stream - top_db) is applied (and written back) on the way.
template <bool CLIP>
__device__ __forceinline__ void block_load(float* row, int n, int lane, float (&x)[4]) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int t = lane + 32 * u;
        x[u] = 0.f;
        if (t < n) x[u] = CLIP ? row[t] : __ldg(row + t);
    }
}

template <int MAXC, bool CLIP>
__device__ __forceinline__ void block_fold(float* row, int n, int cls, float thr, int lane, float (&x)[4],
                                           double (&s)[MAXC], double& q, unsigned& bad) {
    float ps = 0.f, pq = 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int t = lane + 32 * u;
        if (CLIP) {
            if (t < n && x[u] < thr) { x[u] = thr; row[t] = thr; }
        }
        ps += x[u];
        pq = fmaf(x[u], x[u], pq);
    }
    if (!isfinite(pq)) {          // rare: redo this lane's values one by one
        ps = 0.f; pq = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float v = x[u];
            if (!isfinite(v)) { v = 0.f; ++bad; }
            if (fabsf(v) > 1e18f) {   // the square would overflow float32: fold this value in float64
                q = fma((double)v, (double)v, q);
#pragma unroll
                for (int k = 0; k < MAXC; ++k) if (k == cls) s[k] += (double)v;
                v = 0.f;
            }
            ps += v;
            pq = fmaf(v, v, pq);
        }
    }
    const double psd = (double)ps;
    q += (double)pq;
#pragma unroll
    for (int k = 0; k < MAXC; ++k) if (k == cls) s[k] += psd;
}

template <int MAXC, bool CLIP>
__global__ void __launch_bounds__(kThreads, HPSS_MOM_MINB)
moments_kernel(float* __restrict__ feat, const int64_t* __restrict__ frame_off,
               const int32_t* __restrict__ block_clip, int n_clips, int64_t total_frames, int64_t chunk_frames,
               int D, const int32_t* __restrict__ clip_class, int n_classes, double* __restrict__ g_sum,
               double* __restrict__ g_sumsq, double* __restrict__ g_count, double* __restrict__ g_nonfinite,
               const uint32_t* __restrict__ clip_max, int rows_per_stream, int n_streams, float top_db) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = (D + 1) / 2;
    const int d0 = blockIdx.y * kWarps + warp;        // first row of this warp; the second is d0 + half
    if (d0 < half) {
        const int d1 = d0 + half;
        const bool two = d1 < D;
        double s0[MAXC], s1[MAXC];
#pragma unroll
        for (int k = 0; k < MAXC; ++k) { s0[k] = 0.0; s1[k] = 0.0; }
        double q0 = 0.0, q1 = 0.0;
        unsigned bad = 0;
        const int st0 = CLIP ? d0 / rows_per_stream : 0;
        const int st1 = CLIP ? (two ? d1 / rows_per_stream : st0) : 0;
        int64_t g0 = (int64_t)blockIdx.x * chunk_frames;
        const int64_t g1 = min(total_frames, g0 + chunk_frames);
        int c = find_clip_hint(frame_off, block_clip, g0);
        int64_t fo = __ldg(frame_off + c), fe = __ldg(frame_off + c + 1);
        // Software pipeline over 128-frame blocks: block B's offsets, class and thresholds are looked up and
        // its eight loads issued before block A is folded, so neither the metadata chain nor the HBM latency
        // of the next block is exposed.
        struct Blk { float* r0; float* r1; int n; int cls; float thr0, thr1; };
#define HPSS_NEXT_BLOCK(B)                                                                                   \
        do {                                                                                                 \
            (B).n = 0; (B).r0 = feat; (B).r1 = feat; (B).cls = 0; (B).thr0 = -INFINITY; (B).thr1 = -INFINITY; \
            if (g0 < g1) {                                                                                   \
                while (fe <= g0) { ++c; fo = fe; fe = __ldg(frame_off + c + 1); }   /* next non-empty clip */ \
                const int T_ = (int)(fe - fo);                                                               \
                (B).n = (int)(min(min(g1, fe), g0 + 128) - g0);                                              \
                (B).cls = __ldg(clip_class + c);                                                             \
                if (CLIP) {                                                                                  \
                    const uint32_t* cm_ = clip_max + (size_t)n_streams * c;                                  \
                    (B).thr0 = ordered_to_float(__ldg(cm_ + st0)) - top_db;                                  \
                    (B).thr1 = ordered_to_float(__ldg(cm_ + st1)) - top_db;                                  \
                }                                                                                            \
                (B).r0 = feat + (int64_t)D * fo + (int64_t)d0 * T_ + (g0 - fo);                              \
                (B).r1 = (B).r0 + (int64_t)half * T_;                                                        \
                g0 += (B).n;                                                                                 \
            }                                                                                                \
        } while (0)
        Blk A;
        float xa0[4], xa1[4];
        HPSS_NEXT_BLOCK(A);
        block_load<CLIP>(A.r0, A.n, lane, xa0);
        block_load<CLIP>(A.r1, two ? A.n : 0, lane, xa1);
        while (A.n > 0) {
            Blk B;
            float xb0[4], xb1[4];
            HPSS_NEXT_BLOCK(B);
            block_load<CLIP>(B.r0, B.n, lane, xb0);
            block_load<CLIP>(B.r1, two ? B.n : 0, lane, xb1);
            block_fold<MAXC, CLIP>(A.r0, A.n, A.cls, A.thr0, lane, xa0, s0, q0, bad);
            block_fold<MAXC, CLIP>(A.r1, two ? A.n : 0, A.cls, A.thr1, lane, xa1, s1, q1, bad);
            A = B;
#pragma unroll
            for (int u = 0; u < 4; ++u) { xa0[u] = xb0[u]; xa1[u] = xb1[u]; }
        }
#undef HPSS_NEXT_BLOCK
        q0 = warp_sum(q0);
        q1 = warp_sum(q1);
#pragma unroll
        for (int k = 0; k < MAXC; ++k) {
            if (k < n_classes) {
                const double t0 = warp_sum(s0[k]), t1 = warp_sum(s1[k]);
                if (lane == 0 && t0 != 0.0) atomicAdd(g_sum + (size_t)k * D + d0, t0);
                if (lane == 0 && two && t1 != 0.0) atomicAdd(g_sum + (size_t)k * D + d1, t1);
            }
        }
        bad = __reduce_add_sync(0xffffffffu, bad);
        if (lane == 0) {
            if (q0 != 0.0) atomicAdd(g_sumsq + d0, q0);
            if (two && q1 != 0.0) atomicAdd(g_sumsq + d1, q1);
            if (bad) atomicAdd(g_nonfinite, (double)bad);
        }
    }
    // frame counts per class (clip granularity), once
    if (blockIdx.y == 0) {
        for (int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x; c < n_clips; c += (int64_t)gridDim.x * kThreads) {
            const double T = (double)(frame_off[c + 1] - frame_off[c]);
            if (T > 0) atomicAdd(g_count + clip_class[c], T);
        }
    }
}

// The same moments for a batch of equal, short clips (every clip T <= 128 frames: the training-segment shape of
// BASELINE.json configs[1]).  Row d of clip c is the T contiguous floats at ((c * D) + d) * T, so a warp that owns
// row d walks the clips with a constant pointer stride and lane predicates that never change; per clip it needs
// the class and (CLIP) one threshold, nothing else.  Four clips (sixteen loads) are in flight per warp.  The
// per-class float64 sums live in shared memory ([class][lane], one private slot per lane: no atomics), the sum
// of squares in a register.
template <int MAXC, bool CLIP>
__global__ void __launch_bounds__(kThreads)
moments_uniform_kernel(float* __restrict__ feat, int n_clips, int T, int clips_per_chunk, int D,
                       const int32_t* __restrict__ clip_class, int n_classes, double* __restrict__ g_sum,
                       double* __restrict__ g_sumsq, double* __restrict__ g_count, double* __restrict__ g_nonfinite,
                       const uint32_t* __restrict__ clip_max, int rows_per_stream, int n_streams, float top_db) {
    __shared__ double s_cls[kWarps][MAXC][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = blockIdx.y * kWarps + warp;
    constexpr int NC = 4;                                    // clips in flight
    if (d < D) {
#pragma unroll
        for (int k = 0; k < MAXC; ++k) s_cls[warp][k][lane] = 0.0;
        double q = 0.0;
        unsigned bad = 0;
        const int stream = CLIP ? d / rows_per_stream : 0;
        const int c0 = blockIdx.x * clips_per_chunk;
        const int c1 = min(n_clips, c0 + clips_per_chunk);
        const size_t pitch = (size_t)D * T;                  // floats between the same row of consecutive clips
        float* row = feat + (size_t)c0 * pitch + (size_t)d * T;
        bool in[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) in[u] = lane + 32 * u < T;
        for (int c = c0; c < c1; c += NC) {
            float x[NC][4];
            int cls[NC];
            float thr[NC];
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                const bool live = c + i < c1;
                cls[i] = live ? __ldg(clip_class + c + i) : -1;
                thr[i] = -INFINITY;
                if (CLIP && live) thr[i] = ordered_to_float(__ldg(clip_max + (size_t)n_streams * (c + i) + stream)) - top_db;
                float* r = row + (size_t)i * pitch;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    x[i][u] = 0.f;
                    if (live && in[u]) x[i][u] = CLIP ? r[lane + 32 * u] : __ldg(r + lane + 32 * u);
                }
            }
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                if (cls[i] < 0) continue;                    // warp-uniform
                float* r = row + (size_t)i * pitch;
                float ps = 0.f, pq = 0.f;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (CLIP) {
                        if (in[u] && x[i][u] < thr[i]) { x[i][u] = thr[i]; r[lane + 32 * u] = thr[i]; }
                    }
                    ps += x[i][u];
                    pq = fmaf(x[i][u], x[i][u], pq);
                }
                if (!isfinite(pq)) {                         // rare: redo this lane's values one by one
                    ps = 0.f; pq = 0.f;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float v = x[i][u];
                        if (!isfinite(v)) { v = 0.f; ++bad; }
                        if (fabsf(v) > 1e18f) {              // the square would overflow float32
                            q = fma((double)v, (double)v, q);
                            s_cls[warp][cls[i]][lane] += (double)v;
                            v = 0.f;
                        }
                        ps += v;
                        pq = fmaf(v, v, pq);
                    }
                }
                q += (double)pq;
                s_cls[warp][cls[i]][lane] += (double)ps;
            }
            row += (size_t)NC * pitch;
        }
        q = warp_sum(q);
#pragma unroll
        for (int k = 0; k < MAXC; ++k) {
            if (k < n_classes) {
                const double t = warp_sum(s_cls[warp][k][lane]);
                if (lane == 0 && t != 0.0) atomicAdd(g_sum + (size_t)k * D + d, t);
            }
        }
        bad = __reduce_add_sync(0xffffffffu, bad);
        if (lane == 0) {
            if (q != 0.0) atomicAdd(g_sumsq + d, q);
            if (bad) atomicAdd(g_nonfinite, (double)bad);
        }
    }
    // frame counts per class (clip granularity), once
    if (blockIdx.y == 0) {
        for (int c = blockIdx.x * kThreads + threadIdx.x; c < n_clips; c += gridDim.x * kThreads)
            atomicAdd(g_count + clip_class[c], (double)T);
    }
}

// (x - mean) / (stdev + eps) in float64, lane = frame
__global__ void __launch_bounds__(kThreads)
scale_kernel(const float* __restrict__ feat, const int64_t* __restrict__ frame_off,
             const int32_t* __restrict__ block_clip, int64_t total_frames,
             int D, const float* __restrict__ mean, const float* __restrict__ stdev, double eps,
             double* __restrict__ out, float* __restrict__ out32) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t gf = (int64_t)blockIdx.x * 32 + lane;
    if (gf >= total_frames) return;
    const int c = find_clip_hint(frame_off, block_clip, gf);
    const int64_t fo = __ldg(frame_off + c);
    const int T = (int)(__ldg(frame_off + c + 1) - fo);
    const int64_t base = (int64_t)D * fo + (gf - fo);
    for (int d = warp; d < D; d += kWarps) {
        const int64_t g = base + (int64_t)d * T;
        if (out32)      // numpy's float32 evaluation of (FV - mean) / stdev: two rounded float32 operations (:612-613)
            out32[g] = __fdiv_rn(__fsub_rn(__ldg(feat + g), __ldg(mean + d)), __ldg(stdev + d));
        else
            out[g] = ((double)__ldg(feat + g) - (double)__ldg(mean + d)) / ((double)__ldg(stdev + d) + eps);
    }
}

// sklearn.preprocessing.StandardScaler(copy=False).fit_transform on one clip's rows:
// float64 two-pass mean / variance (ddof=0), constant rows keep scale 1, then the in-place float32
// updates X -= float32(mean); X /= float32(scale) (what sklearn 1.9 does for float32 input).
// One warp per (clip, row) line.
__global__ void __launch_bounds__(kThreads)
row_standardize_kernel(float* __restrict__ feat, const int64_t* __restrict__ frame_off, int n_clips, int D) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_lines = (int64_t)n_clips * D;
    for (int64_t line = (int64_t)blockIdx.x * kWarps + warp; line < n_lines; line += (int64_t)gridDim.x * kWarps) {
        const int c = (int)(line / D);
        const int d = (int)(line - (int64_t)c * D);
        const int64_t fo = __ldg(frame_off + c);
        const int T = (int)(__ldg(frame_off + c + 1) - fo);
        if (T <= 0) continue;
        float* x = feat + (int64_t)D * fo + (int64_t)d * T;
        double s = 0.0;
        for (int t = lane; t < T; t += 32) s += (double)x[t];
        const double mean = warp_sum(s) / (double)T;
        double v = 0.0;
        for (int t = lane; t < T; t += 32) {
            const double dlt = (double)x[t] - mean;
            v += dlt * dlt;
        }
        const double var = warp_sum(v) / (double)T;
        const double eps = 2.220446049250313e-16;
        const double ub = (double)T * eps * var + ((double)T * mean * eps) * ((double)T * mean * eps);
        const double scale = (var <= ub) ? 1.0 : sqrt(var);
        const float mean32 = (float)mean, scale32 = (float)scale;   // X -= mean.astype(f32); X /= scale.astype(f32)
        for (int t = lane; t < T; t += 32) x[t] = __fdiv_rn(__fsub_rn(x[t], mean32), scale32);
    }
}

// The same for lines that fit the registers of a warp (T <= 32 * NV): every value is loaded once, mean, variance and
// the update come from registers (same per-lane summation order as the kernel above: bit-identical results), one
// read and one write of the features instead of three reads and one write.
template <int NV>
__global__ void __launch_bounds__(kThreads)
row_standardize_cached_kernel(float* __restrict__ feat, const int64_t* __restrict__ frame_off, int n_clips, int D) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_lines = (int64_t)n_clips * D;
    for (int64_t line = (int64_t)blockIdx.x * kWarps + warp; line < n_lines; line += (int64_t)gridDim.x * kWarps) {
        const int c = (int)(line / D);
        const int d = (int)(line - (int64_t)c * D);
        const int64_t fo = __ldg(frame_off + c);
        const int T = (int)(__ldg(frame_off + c + 1) - fo);
        if (T <= 0) continue;
        float* x = feat + (int64_t)D * fo + (int64_t)d * T;
        float v[NV];
#pragma unroll
        for (int u = 0; u < NV; ++u) v[u] = (lane + 32 * u < T) ? x[lane + 32 * u] : 0.f;
        double s = 0.0;
#pragma unroll
        for (int u = 0; u < NV; ++u) if (lane + 32 * u < T) s += (double)v[u];
        const double mean = warp_sum(s) / (double)T;
        double q = 0.0;
#pragma unroll
        for (int u = 0; u < NV; ++u) {
            if (lane + 32 * u < T) {
                const double dlt = (double)v[u] - mean;
                q += dlt * dlt;
            }
        }
        const double var = warp_sum(q) / (double)T;
        const double eps = 2.220446049250313e-16;
        const double ub = (double)T * eps * var + ((double)T * mean * eps) * ((double)T * mean * eps);
        const double scale = (var <= ub) ? 1.0 : sqrt(var);
        const float mean32 = (float)mean, scale32 = (float)scale;
#pragma unroll
        for (int u = 0; u < NV; ++u)
            if (lane + 32 * u < T) x[lane + 32 * u] = __fdiv_rn(__fsub_rn(v[u], mean32), scale32);
    }
}

// patches[p, d, w] = feat[d, start_p + w] as float64 (tools.pyx:21-38)
__global__ void __launch_bounds__(kThreads)
patches_kernel(const float* __restrict__ feat, int D, int64_t T, int W, int shift, int64_t n_patches,
               double* __restrict__ out) {
    const int64_t total = n_patches * D * (int64_t)W;
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
        const int w = (int)(i % W);
        const int64_t pd = i / W;
        const int d = (int)(pd % D);
        const int64_t p = pd / D;
        int64_t start = p * shift;
        int64_t end = start + W;
        if (end > T) end = T;
        if (end - start < W) start = end - W;
        out[i] = (double)__ldg(feat + (int64_t)d * T + start + w);
    }
}

// flags[c * D + d] = 1 when row d of clip c holds a non-finite value (get_data_stats drops such rows per file,
// lib/preprocessing.py:507-508); one warp per (clip, row) line
__global__ void __launch_bounds__(kThreads)
row_nonfinite_kernel(const float* __restrict__ feat, const int64_t* __restrict__ frame_off, int n_clips, int D,
                     uint8_t* __restrict__ flags) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_lines = (int64_t)n_clips * D;
    for (int64_t line = (int64_t)blockIdx.x * kWarps + warp; line < n_lines; line += (int64_t)gridDim.x * kWarps) {
        const int c = (int)(line / D);
        const int d = (int)(line - (int64_t)c * D);
        const int64_t fo = __ldg(frame_off + c);
        const int T = (int)(__ldg(frame_off + c + 1) - fo);
        const float* x = feat + (int64_t)D * fo + (int64_t)d * T;
        bool bad = false;
        for (int t = lane; t < T; t += 32) bad |= !isfinite(__ldg(x + t));
        bad = __any_sync(0xffffffffu, bad);
        if (lane == 0) flags[line] = bad ? 1 : 0;
    }
}

// N1: model-ready patch tensor.  Patch p of clip c covers frames [p*shift, p*shift + W) of the clip tiled along time
// until it is longer than W (lib/preprocessing.py:139-142: index t mod T_c), rows [row0, row0 + n_rows) of the
// (D, T_c) featuregram.  out[(patch, r, w)] (CNN layout, lib/proposed_architectures.py:451) or, TIME_MAJOR,
// out[(patch, w, r)] (the transposed layout the TCNs take, Proposed_Work_Results.py:235-236).  One CTA = one 32 x 32
// (row, frame) tile of one patch: reads are coalesced along frames, writes along the innermost output axis (through a
// shared-memory transpose when that is the row axis).
template <typename IN, typename OUT, bool TIME_MAJOR>
__global__ void __launch_bounds__(kThreads)
patch_tensor_kernel(const IN* __restrict__ feat, const int64_t* __restrict__ frame_off,
                    const int64_t* __restrict__ patch_off, int n_clips, int D, int row0, int n_rows, int W, int shift,
                    OUT* __restrict__ out) {
    __shared__ OUT tile[32][33];
    __shared__ int s_clip;
    const int64_t p = blockIdx.x;                                  // patches on the x axis (no 65535 limit)
    if (threadIdx.x == 0) s_clip = find_clip(patch_off, n_clips, p);
    __syncthreads();
    const int c = s_clip;
    const int64_t fo = __ldg(frame_off + c);
    const int T = (int)(__ldg(frame_off + c + 1) - fo);
    const int64_t start = (p - __ldg(patch_off + c)) * shift;
    const IN* base = feat + (int64_t)D * fo + (int64_t)row0 * T;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;        // 32 x 8
    const int r0 = blockIdx.y * 32;                                // one CTA = 32 rows x the whole patch width
    OUT* op = out + (int64_t)p * n_rows * W;
    for (int w0 = 0; w0 < W; w0 += 32) {
        const int w = w0 + tx;
        const int t = w < W ? (int)((start + w) % T) : 0;
        OUT v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = r0 + ty + 8 * i;
            v[i] = (OUT)0;
            if (r < n_rows && w < W) v[i] = (OUT)__ldg(base + (int64_t)r * T + t);
        }
        if (!TIME_MAJOR) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = r0 + ty + 8 * i;
                if (r < n_rows && w < W) op[(int64_t)r * W + w] = v[i];
            }
        } else {
            if (w0) __syncthreads();                               // the previous tile has been read
#pragma unroll
            for (int i = 0; i < 4; ++i) tile[ty + 8 * i][tx] = v[i];
            __syncthreads();
            const int r = r0 + tx;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ww = w0 + ty + 8 * i;
                if (r < n_rows && ww < W) op[(int64_t)ww * n_rows + r] = tile[tx][ty + 8 * i];
            }
        }
    }
}

// get_data_statistics (lib/cython_impl/tools.pyx:169-211): per patch, mean / variance (ddof = 0) / scipy.stats.skew /
// scipy.stats.kurtosis (biased, Fisher) along one axis of a (N, A, B) float64 array.  ALONG_A: reduce over A ->
// (N, B) (the reference's axis=0, "percussive"); else reduce over B -> (N, A) (axis=1, "harmonic").  One warp per
// output value, two passes (mean, then central moments) like numpy / scipy.
template <bool ALONG_A>
__global__ void __launch_bounds__(kThreads)
patch_stats_kernel(const double* __restrict__ x, int64_t N, int A, int B, int stat, double* __restrict__ out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_red = ALONG_A ? A : B, n_keep = ALONG_A ? B : A;
    const int64_t total = N * n_keep;
    for (int64_t o = (int64_t)blockIdx.x * kWarps + warp; o < total; o += (int64_t)gridDim.x * kWarps) {
        const int64_t n = o / n_keep;
        const int j = (int)(o - n * n_keep);
        const double* p = x + n * (int64_t)A * B + (ALONG_A ? j : (int64_t)j * B);
        const int64_t stride = ALONG_A ? B : 1;
        double s = 0.0;
        for (int i = lane; i < n_red; i += 32) s += p[(int64_t)i * stride];
        const double mean = warp_sum(s) / (double)n_red;
        double m2 = 0.0, m3 = 0.0, m4 = 0.0;
        for (int i = lane; i < n_red; i += 32) {
            const double d = p[(int64_t)i * stride] - mean, d2 = d * d;
            m2 += d2; m3 += d2 * d; m4 += d2 * d2;
        }
        m2 = warp_sum(m2) / (double)n_red;
        m3 = warp_sum(m3) / (double)n_red;
        m4 = warp_sum(m4) / (double)n_red;
        if (lane == 0) {
            // scipy: NaN where the variance is lost to rounding (m2 <= (eps * mean)^2)
            const double eps = 2.220446049250313e-16;
            const bool zero = m2 <= (eps * mean) * (eps * mean);
            double r;
            if (stat == 0) r = mean;
            else if (stat == 1) r = m2;
            else if (stat == 2) r = zero ? nan("") : m3 / (m2 * sqrt(m2));
            else r = zero ? nan("") : m4 / (m2 * m2) - 3.0;
            out[o] = r;
        }
    }
}

}  // namespace

int launch_row_nonfinite(hpss_ctx* ctx, const hpss_batch* b, const float* feat, int D, uint8_t* flags, cudaStream_t st) {
    const int64_t n_lines = (int64_t)b->n_clips * D;
    if (n_lines == 0) return HPSS_OK;
    int64_t grid = (n_lines + kWarps - 1) / kWarps;
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    if (grid > cap) grid = cap;
    row_nonfinite_kernel<<<(unsigned)grid, kThreads, 0, st>>>(feat, b->d_frame_off, b->n_clips, D, flags);
    HPSS_LAUNCHED("row_nonfinite_kernel");
    return HPSS_OK;
}

int launch_patch_tensor(const hpss_batch* b, const void* feat, int in_f64, const int64_t* d_patch_off, int64_t n_patches,
                        int D, int row0, int n_rows, int W, int shift, int time_major, int out_f64, void* out,
                        cudaStream_t st) {
    if (n_patches == 0) return HPSS_OK;
    if (n_patches > 0x7fffffffLL) { set_error("patch_tensor: too many patches"); return HPSS_ERR_UNSUPPORTED; }
    dim3 grid((unsigned)n_patches, (unsigned)((n_rows + 31) / 32), 1);
#define HPSS_PT(IN, OUT, TM) patch_tensor_kernel<IN, OUT, TM><<<grid, kThreads, 0, st>>>((const IN*)feat, b->d_frame_off, d_patch_off, b->n_clips, D, row0, n_rows, W, shift, (OUT*)out)
    if (in_f64) {       // float64 featuregrams (the output of the Cython scale_data): float64 patches, exact copies
        if (!out_f64) { set_error("patch_tensor: float64 input needs float64 output"); return HPSS_ERR_INVALID; }
        if (time_major) HPSS_PT(double, double, true); else HPSS_PT(double, double, false);
    } else if (out_f64) { if (time_major) HPSS_PT(float, double, true); else HPSS_PT(float, double, false); }
    else { if (time_major) HPSS_PT(float, float, true); else HPSS_PT(float, float, false); }
#undef HPSS_PT
    HPSS_LAUNCHED("patch_tensor_kernel");
    return HPSS_OK;
}

int launch_patch_stats(hpss_ctx* ctx, const double* x, int64_t N, int A, int B, int stat, int along_a, double* out,
                       cudaStream_t st) {
    const int64_t total = N * (along_a ? B : A);
    if (total == 0) return HPSS_OK;
    int64_t grid = (total + kWarps - 1) / kWarps;
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (grid > cap) grid = cap;
    if (along_a) patch_stats_kernel<true><<<(unsigned)grid, kThreads, 0, st>>>(x, N, A, B, stat, out);
    else patch_stats_kernel<false><<<(unsigned)grid, kThreads, 0, st>>>(x, N, A, B, stat, out);
    HPSS_LAUNCHED("patch_stats_kernel");
    return HPSS_OK;
}

static int launch_moments_impl(hpss_ctx* ctx, const hpss_batch* b, float* feat, int D, const int32_t* d_class,
                               int n_classes, double* sum, double* sumsq, double* count, double* nonfinite,
                               const uint32_t* clip_max, int rows_per_stream, int n_streams, float top_db,
                               cudaStream_t st) {
    const int64_t total = b->frame_off[b->n_clips];
    if (total == 0) return HPSS_OK;
    if (n_classes > 8) {
        set_error("moments: at most 8 classes are supported (got %d)", n_classes);
        return HPSS_ERR_UNSUPPORTED;
    }
    if (b->uniform_frames > 0 && b->uniform_frames <= 128 && !knobs().no_uniform_moments) {
        // equal short clips: constant strides, no per-clip table lookups (moments_uniform_kernel)
        const int T = (int)b->uniform_frames;
        const int rg = (D + kWarps - 1) / kWarps;
        const int ctas_per_sm = knobs().mom_ctas;
        int want = (ctx->sm_count * ctas_per_sm + rg - 1) / rg;      // ~16 CTAs per SM in total (48: 0.110 ms, 16: 0.103 ms, 8: 0.107 ms on configs[1])
        int per = (b->n_clips + want - 1) / want;
        per = std::max(8, (per + 3) / 4 * 4);
        dim3 grid((unsigned)((b->n_clips + per - 1) / per), (unsigned)rg);
#define HPSS_UMOMENTS_LAUNCH(MAXC, CLIP)                                                                              \
        moments_uniform_kernel<MAXC, CLIP><<<grid, kThreads, 0, st>>>(feat, b->n_clips, T, per, D, d_class, n_classes, \
                                                                      sum, sumsq, count, nonfinite, clip_max,          \
                                                                      rows_per_stream, n_streams, top_db)
        if (clip_max) {
            if (n_classes <= 4) HPSS_UMOMENTS_LAUNCH(4, true); else HPSS_UMOMENTS_LAUNCH(8, true);
        } else {
            if (n_classes <= 4) HPSS_UMOMENTS_LAUNCH(4, false); else HPSS_UMOMENTS_LAUNCH(8, false);
        }
#undef HPSS_UMOMENTS_LAUNCH
        HPSS_LAUNCHED("moments_uniform_kernel");
        return HPSS_OK;
    }
    // frame chunks: multiples of 32 frames, enough CTAs for ~8 waves of 8 resident CTAs per SM
    const int row_groups = ((D + 1) / 2 + kWarps - 1) / kWarps;   // a warp owns two rows
    int64_t want_chunks = ((int64_t)ctx->sm_count * 64 + row_groups - 1) / row_groups;
    int64_t chunk = (total + want_chunks - 1) / want_chunks;
    chunk = (chunk + 31) / 32 * 32;
    if (chunk < 1024) chunk = 1024;
    const int64_t n_chunks = (total + chunk - 1) / chunk;
    dim3 grid((unsigned)n_chunks, (unsigned)row_groups);
#define HPSS_MOMENTS_LAUNCH(MAXC, CLIP)                                                                              \
    moments_kernel<MAXC, CLIP><<<grid, kThreads, 0, st>>>(feat, b->d_frame_off, b->d_block_clip, b->n_clips, total,  \
                                                          chunk, D, d_class, n_classes, sum, sumsq, count, nonfinite, \
                                                          clip_max, rows_per_stream, n_streams, top_db)
    if (clip_max) {
        if (n_classes <= 4) HPSS_MOMENTS_LAUNCH(4, true); else HPSS_MOMENTS_LAUNCH(8, true);
    } else {
        if (n_classes <= 4) HPSS_MOMENTS_LAUNCH(4, false); else HPSS_MOMENTS_LAUNCH(8, false);
    }
#undef HPSS_MOMENTS_LAUNCH
    HPSS_LAUNCHED("moments_kernel");
    return HPSS_OK;
}

int launch_moments(hpss_ctx* ctx, const hpss_batch* b, const float* feat, int D, const int32_t* d_class,
                   int n_classes, double* sum, double* sumsq, double* count, double* nonfinite, cudaStream_t st) {
    return launch_moments_impl(ctx, b, const_cast<float*>(feat), D, d_class, n_classes, sum, sumsq, count, nonfinite,
                               nullptr, 1, 1, 0.f, st);
}

int launch_topdb_moments(hpss_ctx* ctx, const hpss_batch* b, float* feat, int rows_per_stream, int n_streams,
                         const uint32_t* clip_max, float top_db, const int32_t* d_class, int n_classes, double* sum,
                         double* sumsq, double* count, double* nonfinite, cudaStream_t st) {
    return launch_moments_impl(ctx, b, feat, rows_per_stream * n_streams, d_class, n_classes, sum, sumsq, count,
                               nonfinite, clip_max, rows_per_stream, n_streams, top_db, st);
}

int launch_scale(hpss_ctx* ctx, const hpss_batch* b, const float* feat, int D, const float* mean,
                 const float* stdev, double eps, double* out, float* out32, cudaStream_t st) {
    (void)ctx;
    const int64_t total = b->frame_off[b->n_clips];
    if (total == 0) return HPSS_OK;
    scale_kernel<<<(unsigned)((total + 31) / 32), kThreads, 0, st>>>(feat, b->d_frame_off, b->d_block_clip, total, D, mean,
                                                                     stdev, eps, out, out32);
    HPSS_LAUNCHED("scale_kernel");
    return HPSS_OK;
}

int launch_row_standardize(hpss_ctx* ctx, const hpss_batch* b, float* feat, int D, cudaStream_t st) {
    const int64_t n_lines = (int64_t)b->n_clips * D;
    if (n_lines == 0) return HPSS_OK;
    int64_t grid = (n_lines + kWarps - 1) / kWarps;
    if (b->max_frames <= 1024) {                      // lines that fit a warp's registers: one read, one write
        const int64_t cap = (int64_t)ctx->sm_count * 64;
        if (grid > cap) grid = cap;
        if (b->max_frames <= 128) row_standardize_cached_kernel<4><<<(unsigned)grid, kThreads, 0, st>>>(feat, b->d_frame_off, b->n_clips, D);
        else row_standardize_cached_kernel<32><<<(unsigned)grid, kThreads, 0, st>>>(feat, b->d_frame_off, b->n_clips, D);
        HPSS_LAUNCHED("row_standardize_cached_kernel");
        return HPSS_OK;
    }
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    if (grid > cap) grid = cap;
    row_standardize_kernel<<<(unsigned)grid, kThreads, 0, st>>>(feat, b->d_frame_off, b->n_clips, D);
    HPSS_LAUNCHED("row_standardize_kernel");
    return HPSS_OK;
}

int launch_patches(const float* feat, int D, int64_t T, int W, int shift, int64_t n_patches, double* out,
                   cudaStream_t st) {
    const int64_t total = n_patches * D * (int64_t)W;
    if (total <= 0) return HPSS_OK;
    int64_t grid = (total + kThreads - 1) / kThreads;
    if (grid > kSMs * 16) grid = kSMs * 16;
    patches_kernel<<<(unsigned)grid, kThreads, 0, st>>>(feat, D, T, W, shift, n_patches, out);
    HPSS_LAUNCHED("patches_kernel");
    return HPSS_OK;
}

}  // namespace hpss
