// K3 / K3b: soft masks + masked spectrogram + banded mel projection + power_to_db.
//
// Replaces, per clip, the tail of librosa.decompose.hpss (util.softmask(power=2,
// split_zeros=True) twice and S*mask), librosa.feature.melspectrogram(S=., n_mels=M)
// (np.dot with the Slaney basis) and librosa.core.power_to_db(.**2)
// (lib/preprocessing.py:408-412, 418-424, 430-434, 440-444 of the reference).
//
// The soft-mask arithmetic uses the IEEE round-to-nearest intrinsics in numpy's operation
// order (no FMA contraction), so H = S*mask_h and P = S*mask_p are bit-identical to the
// reference for identical (S, harm, perc).  The mel basis is banded (triangles): only the
// [first,last) non-zero columns of each filter are visited, in increasing f, fp32 FMA.
//
// One CTA owns 32 consecutive frames of the batch (lane = frame, so every global access is
// a coalesced 128-byte row segment); the 8 warps split the frequency rows (mask phase) and
// the mel filters (projection phase); masked values are staged through shared memory in
// chunks of 64 frequency rows.
#include <stdlib.h>

#include "maskmath.cuh"

namespace hpss {

namespace {

#ifndef HPSS_K3_U
#define HPSS_K3_U 4      // frequency rows in flight per warp in the sweep kernel
#endif
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kFCsplit = 64;    // frequency rows staged per chunk when a whole column does not fit
constexpr int kFCwhole = 264;   // up to this many rows the whole column is staged at once

struct FrameLane {
    bool valid;
    int clip;
    int T;
    int64_t in_base;    // + f*T
    int64_t out_base;   // + row*T
};

__device__ __forceinline__ FrameLane frame_lane(const int64_t* __restrict__ frame_off, const int32_t* __restrict__ block_clip, int64_t total,
                                                int64_t gf, int rows_in, int rows_out) {
    FrameLane fl;
    fl.valid = gf < total;
    fl.clip = 0; fl.T = 1; fl.in_base = 0; fl.out_base = 0;
    if (fl.valid) {
        const int c = find_clip_hint(frame_off, block_clip, gf);
        const int64_t fo = __ldg(frame_off + c);
        fl.clip = c;
        fl.T = (int)(__ldg(frame_off + c + 1) - fo);
        fl.in_base = (int64_t)rows_in * fo + (gf - fo);
        fl.out_base = (int64_t)rows_out * fo + (gf - fo);
    }
    return fl;
}

template <bool HPSS_MODE>
__global__ void __launch_bounds__(kThreads)
mask_mel_kernel(const float* __restrict__ S, const float* __restrict__ harm, const float* __restrict__ perc,
                const int64_t* __restrict__ frame_off, const int32_t* __restrict__ block_clip, int64_t total_frames, int rows,
                const float* __restrict__ mel, const int2* __restrict__ band, int n_mels, int pre_square,
                int log_power, float amin, float* __restrict__ out, uint32_t* __restrict__ clip_max, int FC) {
    constexpr int NS = HPSS_MODE ? 2 : 1;
    extern __shared__ float smem[];
    __shared__ float s_max[NS][kWarps][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool project = mel != nullptr;
    const int rows_out = NS * (project ? n_mels : rows);
    const FrameLane fl = frame_lane(frame_off, block_clip, total_frames, (int64_t)blockIdx.x * 32 + lane, rows, rows_out);

    float vmax[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) vmax[s] = -INFINITY;   // only constant indices below (stays in registers)

    if (!project) {
        // identity projection: [H; P] (or plain S) rows, optional power_to_db
        for (int f = warp; f < rows; f += kWarps) {
            if (!fl.valid) continue;
            const int64_t gi = fl.in_base + (int64_t)f * fl.T;
            float sv = __ldg(S + gi);
            if (HPSS_MODE) {
                float H, P;
                softmask_apply(sv, __ldg(harm + gi), __ldg(perc + gi), H, P);
                const float a = post_value(H, log_power, amin), b = post_value(P, log_power, amin);
                out[fl.out_base + (int64_t)f * fl.T] = a;
                out[fl.out_base + (int64_t)(rows + f) * fl.T] = b;
                vmax[0] = fmaxf(vmax[0], a);
                vmax[NS - 1] = fmaxf(vmax[NS - 1], b);
            } else {
                if (pre_square) sv = __fmul_rn(sv, sv);
                const float a = post_value(sv, log_power, amin);
                out[fl.out_base + (int64_t)f * fl.T] = a;
                vmax[0] = fmaxf(vmax[0], a);
            }
        }
    } else {
        float* Hc = smem;                       // [FC][32]
        float* Pc = Hc + FC * 32;               // [FC][32]   (HPSS mode only)
        if (FC >= rows) {
            // whole column staged at once: mask phase, then every filter accumulates in registers.
            // Four rows per iteration with all twelve loads issued first (memory-level parallelism).
            constexpr int U = 4;
            for (int fb = warp; fb < rows; fb += U * kWarps) {
                float sv[U], hv[U], pv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int f = fb + u * kWarps;
                    sv[u] = 0.f; hv[u] = 0.f; pv[u] = 0.f;
                    if (fl.valid && f < rows) {
                        const int64_t gi = fl.in_base + (int64_t)f * fl.T;
                        sv[u] = __ldg(S + gi);
                        if (HPSS_MODE) { hv[u] = __ldg(harm + gi); pv[u] = __ldg(perc + gi); }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int f = fb + u * kWarps;
                    if (f < rows) {
                        float H = 0.f, P = 0.f;
                        if (HPSS_MODE) softmask_apply(sv[u], hv[u], pv[u], H, P);
                        else H = pre_square ? __fmul_rn(sv[u], sv[u]) : sv[u];
                        Hc[f * 32 + lane] = H;
                        if (HPSS_MODE) Pc[f * 32 + lane] = P;
                    }
                }
            }
            __syncthreads();
            for (int m = warp; m < n_mels; m += kWarps) {
                const int2 bd = __ldg(band + m);
                float sh = 0.f, sp = 0.f;
                const float* w = mel + (size_t)m * rows;
                for (int f = bd.x; f < bd.y; ++f) {
                    const float wf = __ldg(w + f);
                    sh = fmaf(wf, Hc[f * 32 + lane], sh);
                    if (HPSS_MODE) sp = fmaf(wf, Pc[f * 32 + lane], sp);
                }
                if (fl.valid) {
                    const float a = post_value(sh, log_power, amin);
                    out[fl.out_base + (int64_t)m * fl.T] = a;
                    vmax[0] = fmaxf(vmax[0], a);
                    if (HPSS_MODE) {
                        const float b = post_value(sp, log_power, amin);
                        out[fl.out_base + (int64_t)(n_mels + m) * fl.T] = b;
                        vmax[NS - 1] = fmaxf(vmax[NS - 1], b);
                    }
                }
            }
        } else {
            float* acc = HPSS_MODE ? Pc + FC * 32 : Pc;   // [NS][n_mels][32]
            for (int i = threadIdx.x; i < NS * n_mels * 32; i += kThreads) acc[i] = 0.f;
            for (int f0 = 0; f0 < rows; f0 += FC) {
                const int fe = min(rows, f0 + FC);
                // phase 1: masked values of this chunk -> shared memory
                for (int f = f0 + warp; f < fe; f += kWarps) {
                    float H = 0.f, P = 0.f;
                    if (fl.valid) {
                        const int64_t gi = fl.in_base + (int64_t)f * fl.T;
                        const float sv = __ldg(S + gi);
                        if (HPSS_MODE) softmask_apply(sv, __ldg(harm + gi), __ldg(perc + gi), H, P);
                        else H = pre_square ? __fmul_rn(sv, sv) : sv;
                    }
                    Hc[(f - f0) * 32 + lane] = H;
                    if (HPSS_MODE) Pc[(f - f0) * 32 + lane] = P;
                }
                __syncthreads();
                // phase 2: banded projection of the chunk
                for (int m = warp; m < n_mels; m += kWarps) {
                    const int2 bd = __ldg(band + m);
                    const int a = max(bd.x, f0), b = min(bd.y, fe);
                    if (a >= b) continue;
                    float sh = 0.f, sp = 0.f;
                    const float* w = mel + (size_t)m * rows;
                    for (int f = a; f < b; ++f) {
                        const float wf = __ldg(w + f);
                        sh = fmaf(wf, Hc[(f - f0) * 32 + lane], sh);
                        if (HPSS_MODE) sp = fmaf(wf, Pc[(f - f0) * 32 + lane], sp);
                    }
                    acc[m * 32 + lane] += sh;
                    if (HPSS_MODE) acc[(n_mels + m) * 32 + lane] += sp;
                }
                __syncthreads();
            }
            for (int r = warp; r < NS * n_mels; r += kWarps) {
                if (!fl.valid) continue;
                const float a = post_value(acc[r * 32 + lane], log_power, amin);
                out[fl.out_base + (int64_t)r * fl.T] = a;
                if (NS == 2 && r >= n_mels) vmax[NS - 1] = fmaxf(vmax[NS - 1], a);
                else vmax[0] = fmaxf(vmax[0], a);
            }
        }
    }

    if (clip_max != nullptr) {
#pragma unroll
        for (int s = 0; s < NS; ++s) s_max[s][warp][lane] = vmax[s];
        __syncthreads();
        if (warp < NS) {
            float v = -INFINITY;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) v = fmaxf(v, s_max[warp][w][lane]);
            publish_max(clip_max, NS, warp, fl.valid, fl.clip, v);
        }
    }
}

// Sweep formulation of the HPSS mel features (the default whenever the basis is "sweepable": bands ordered,
// at most two filters overlapping at any frequency row -- true for every Slaney basis the reference builds).
// One warp owns 32 consecutive frames of the batch (lane = frame) and walks f = 0 .. rows-1 once: the three
// loads per row are coalesced 128-byte segments, U rows are in flight at a time, the masked values go
// straight into the two running sums per stream of the filters currently open and a finished filter is
// emitted (square, log, store, running max) when f passes its upper edge.  No shared memory, no CTA
// barrier, fully independent warps: occupancy (not a tile ring) hides the HBM latency.  Accumulation order
// (f ascending, fmaf) is the one of mask_mel_kernel's whole-column path, so both give the same bits.
template <int U>
__global__ void __launch_bounds__(kThreads)
mask_mel_sweep_kernel(const float* __restrict__ S, const float* __restrict__ harm, const float* __restrict__ perc,
                      const int64_t* __restrict__ frame_off, const int32_t* __restrict__ block_clip,
                      int64_t total_frames, int rows, const int4* __restrict__ sweep, int n_mels, int log_power,
                      float amin, float* __restrict__ out, uint32_t* __restrict__ clip_max) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t g0 = ((int64_t)blockIdx.x * kWarps + warp) * 32;
    if (g0 >= total_frames) return;
    const FrameLane fl = frame_lane(frame_off, block_clip, total_frames, g0 + lane, rows, 2 * n_mels);
    const int64_t T = fl.T;
    const float* sp = S + fl.in_base;
    const float* hp = harm + fl.in_base;
    const float* pp = perc + fl.in_base;
    float* oh = out + fl.out_base;
    float* op = oh + (int64_t)n_mels * T;
    float aH = 0.f, aP = 0.f, bH = 0.f, bP = 0.f;     // running sums of filters cur, cur + 1
    float vmaxH = -INFINITY, vmaxP = -INFINITY;
    int cur = 0;

    auto emit = [&]() {
        const float vH = post_value(aH, log_power, amin);
        const float vP = post_value(aP, log_power, amin);
        if (fl.valid) { *oh = vH; *op = vP; }
        oh += T; op += T;
        vmaxH = fmaxf(vmaxH, vH);
        vmaxP = fmaxf(vmaxP, vP);
        aH = bH; aP = bP; bH = 0.f; bP = 0.f;
        ++cur;
    };
    auto row = [&](int f, float sv, float hv, float pv) {
        float H, P;
        softmask_apply(sv, hv, pv, H, P);
        const int4 e = __ldg(sweep + f);
#pragma unroll 1
        while (cur < e.x) emit();                      // warp-uniform: the table does not depend on the lane
        const float wA = __int_as_float(e.y), wB = __int_as_float(e.z);
        aH = fmaf(wA, H, aH);
        aP = fmaf(wA, P, aP);
        bH = fmaf(wB, H, bH);
        bP = fmaf(wB, P, bP);
    };

    int f0 = 0;
    for (; f0 + U <= rows; f0 += U) {
        float sv[U], hv[U], pv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            sv[u] = 0.f; hv[u] = 0.f; pv[u] = 0.f;
            if (fl.valid) {
                sv[u] = __ldg(sp + u * T);
                hv[u] = __ldg(hp + u * T);
                pv[u] = __ldg(pp + u * T);
            }
        }
        sp += U * T; hp += U * T; pp += U * T;
#pragma unroll
        for (int u = 0; u < U; ++u) row(f0 + u, sv[u], hv[u], pv[u]);
    }
    for (; f0 < rows; ++f0) {
        float sv = 0.f, hv = 0.f, pv = 0.f;
        if (fl.valid) { sv = __ldg(sp); hv = __ldg(hp); pv = __ldg(pp); }
        sp += T; hp += T; pp += T;
        row(f0, sv, hv, pv);
    }
#pragma unroll 1
    while (cur < n_mels) emit();                       // filters above the last frequency row
    if (clip_max != nullptr) {
        publish_max(clip_max, 2, 0, fl.valid, fl.clip, vmaxH);
        publish_max(clip_max, 2, 1, fl.valid, fl.clip, vmaxP);
    }
}

// Same sweep with the split tables of the mel plan (emission counts, 4 bits per row + the two weights per row):
// no load sits in front of a branch, the U soft masks of an iteration are one basic block of interleaved
// chains (softmask_batch) and power_to_db is resolved at compile time.  Bit-identical to the kernel above.
template <int U, int LOGP>
__global__ void __launch_bounds__(kThreads)
mask_mel_sweep2_kernel(const float* __restrict__ S, const float* __restrict__ harm, const float* __restrict__ perc,
                       const int64_t* __restrict__ frame_off, const int32_t* __restrict__ block_clip,
                       int64_t total_frames, int rows, const uint32_t* __restrict__ emit4,
                       const float2* __restrict__ sweep_w, int n_mels, float amin, float* __restrict__ out,
                       uint32_t* __restrict__ clip_max) {
    static_assert(8 % U == 0, "U rows never straddle a word of the emission table");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t g0 = ((int64_t)blockIdx.x * kWarps + warp) * 32;
    if (g0 >= total_frames) return;
    const FrameLane fl = frame_lane(frame_off, block_clip, total_frames, g0 + lane, rows, 2 * n_mels);
    const int T = fl.T;
    const float* sp = S + fl.in_base;
    const float* hp = harm + fl.in_base;
    const float* pp = perc + fl.in_base;
    float* oh = out + fl.out_base;
    float* op = oh + (int64_t)n_mels * T;
    float aH = 0.f, aP = 0.f, bH = 0.f, bP = 0.f;     // running sums of filters cur, cur + 1
    float vmaxH = -INFINITY, vmaxP = -INFINITY;
    int cur = 0;
    auto emit = [&]() {
        const float vH = post_value(aH, LOGP, amin);
        const float vP = post_value(aP, LOGP, amin);
        if (fl.valid) { *oh = vH; *op = vP; }
        oh += T; op += T;
        vmaxH = fmaxf(vmaxH, vH);
        vmaxP = fmaxf(vmaxP, vP);
        aH = bH; aP = bP; bH = 0.f; bP = 0.f;
        ++cur;
    };
    const int rows_u = (rows + U - 1) / U * U;         // the tables are zero padded to a multiple of 64 rows
    const uint32_t T4 = 4u * (uint32_t)T;              // row pitch in bytes: one IMAD.WIDE.U32 per address (FMA pipe)
    auto rowp = [&](const float* p, uint32_t f) {
        return reinterpret_cast<const float*>(reinterpret_cast<const char*>(p) + (uint64_t)f * T4);
    };
#pragma unroll 1
    for (int f0 = 0; f0 < rows_u; f0 += U) {
        float sv[U], hv[U], pv[U];
        float2 w[U];
        const uint32_t em = __ldg(emit4 + (f0 >> 3)) >> (4 * (f0 & 7));
#pragma unroll
        for (int u = 0; u < U; ++u) {
            sv[u] = 0.f; hv[u] = 0.f; pv[u] = 0.f;
            if (fl.valid && f0 + u < rows) {
                sv[u] = __ldg(rowp(sp, f0 + u));
                hv[u] = __ldg(rowp(hp, f0 + u));
                pv[u] = __ldg(rowp(pp, f0 + u));
            }
            w[u] = __ldg(sweep_w + f0 + u);
        }
        float H[U], P[U];
        softmask_batch<U>(sv, hv, pv, H, P);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int n = (int)((em >> (4 * u)) & 15u);          // warp-uniform
#pragma unroll 1
            for (; n > 0; --n) emit();
            aH = fmaf(w[u].x, H[u], aH);
            aP = fmaf(w[u].x, P[u], aP);
            bH = fmaf(w[u].y, H[u], bH);
            bP = fmaf(w[u].y, P[u], bP);
        }
    }
#pragma unroll 1
    while (cur < n_mels) emit();                       // filters above the last frequency row
    if (clip_max != nullptr) {
        publish_max(clip_max, 2, 0, fl.valid, fl.clip, vmaxH);
        publish_max(clip_max, 2, 1, fl.valid, fl.clip, vmaxP);
    }
}

__global__ void __launch_bounds__(kThreads)
topdb_kernel(float* __restrict__ out, const int64_t* __restrict__ frame_off, const int32_t* __restrict__ block_clip, int64_t total_frames,
             int rows_per_stream, int n_streams, const uint32_t* __restrict__ clip_max, float top_db) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rows = rows_per_stream * n_streams;
    const FrameLane fl = frame_lane(frame_off, block_clip, total_frames, (int64_t)blockIdx.x * 32 + lane, rows, rows);
    if (!fl.valid) return;
    for (int s = 0; s < n_streams; ++s) {
        const float thr = ordered_to_float(__ldg(clip_max + (size_t)n_streams * fl.clip + s)) - top_db;
        for (int r = warp; r < rows_per_stream; r += kWarps) {
            float* p = out + fl.out_base + (int64_t)(s * rows_per_stream + r) * fl.T;
            *p = fmaxf(*p, thr);
        }
    }
}

__global__ void mel_band_kernel(const float* __restrict__ mel, int n_mels, int rows, int2* __restrict__ band) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_mels) return;
    int first = rows, last = 0;
    for (int f = 0; f < rows; ++f) {
        if (mel[(size_t)m * rows + f] != 0.f) {
            if (f < first) first = f;
            last = f + 1;
        }
    }
    if (first >= last) { first = 0; last = 0; }
    band[m] = make_int2(first, last);
}

}  // namespace

int launch_mel_bands(const float* mel, int n_mels, int rows, int2* band, cudaStream_t st) {
    mel_band_kernel<<<(n_mels + 127) / 128, 128, 0, st>>>(mel, n_mels, rows, band);
    HPSS_LAUNCHED("mel_band_kernel");
    return HPSS_OK;
}

int launch_mask_mel(hpss_ctx* ctx, const hpss_batch* b, const float* S, const float* harm, const float* perc,
                    int rows, const float* mel, const int2* band, const int4* sweep, const uint32_t* emit4,
                    const float2* sweep_w, int n_mels, int pre_square, int log_power, float amin, float* out,
                    uint32_t* clip_max, cudaStream_t st) {
    const int64_t total = b->frame_off[b->n_clips];
    const bool hpss_mode = harm != nullptr;
    const int ns = hpss_mode ? 2 : 1;
    if (clip_max) HPSS_CUDA(cudaMemsetAsync(clip_max, 0, sizeof(uint32_t) * (size_t)ns * b->n_clips, st));
    if (total == 0) return HPSS_OK;
    if (hpss_mode && mel && emit4 && sweep_w && (log_power == 0 || log_power == 1) && !knobs().sweep1) {
        const int64_t n_warps = (total + 31) / 32;
        const unsigned grid = (unsigned)((n_warps + kWarps - 1) / kWarps);
        if (log_power)
            mask_mel_sweep2_kernel<HPSS_K3_U, 1><<<grid, kThreads, 0, st>>>(S, harm, perc, b->d_frame_off, b->d_block_clip, total, rows,
                                                                    emit4, sweep_w, n_mels, amin, out, clip_max);
        else
            mask_mel_sweep2_kernel<HPSS_K3_U, 0><<<grid, kThreads, 0, st>>>(S, harm, perc, b->d_frame_off, b->d_block_clip, total, rows,
                                                                    emit4, sweep_w, n_mels, amin, out, clip_max);
        HPSS_LAUNCHED("mask_mel_sweep2_kernel");
        return HPSS_OK;
    }
    if (hpss_mode && mel && sweep) {
        const int64_t n_warps = (total + 31) / 32;
        const unsigned grid = (unsigned)((n_warps + kWarps - 1) / kWarps);
        const int U = knobs().sweep_u;   // development knob
#define HPSS_SWEEP_LAUNCH(UU)                                                                                       \
        mask_mel_sweep_kernel<UU><<<grid, kThreads, 0, st>>>(S, harm, perc, b->d_frame_off, b->d_block_clip, total,  \
                                                             rows, sweep, n_mels, log_power, amin, out, clip_max)
        if (U == 2) HPSS_SWEEP_LAUNCH(2); else if (U == 8) HPSS_SWEEP_LAUNCH(8); else if (U == 6) HPSS_SWEEP_LAUNCH(6); else HPSS_SWEEP_LAUNCH(4);
#undef HPSS_SWEEP_LAUNCH
        HPSS_LAUNCHED("mask_mel_sweep_kernel");
        return HPSS_OK;
    }
    size_t smem = 0;
    int FC = 0;
    if (mel) {
        if (rows <= kFCwhole) { FC = rows; smem = (size_t)ns * FC * 32 * sizeof(float); }
        else { FC = kFCsplit; smem = ((size_t)ns * FC * 32 + (size_t)ns * n_mels * 32) * sizeof(float); }
    }
    if (smem > (size_t)ctx->max_smem_optin) {
        set_error("n_mels=%d needs %zu bytes of shared memory (max %d)", n_mels, smem, ctx->max_smem_optin);
        return HPSS_ERR_UNSUPPORTED;
    }
    const unsigned grid = (unsigned)((total + 31) / 32);
    if (hpss_mode) {
        HPSS_CUDA(cudaFuncSetAttribute(mask_mel_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mask_mel_kernel<true><<<grid, kThreads, smem, st>>>(S, harm, perc, b->d_frame_off, b->d_block_clip, total, rows,
                                                            mel, band, n_mels, pre_square, log_power, amin, out,
                                                            clip_max, FC);
    } else {
        HPSS_CUDA(cudaFuncSetAttribute(mask_mel_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mask_mel_kernel<false><<<grid, kThreads, smem, st>>>(S, nullptr, nullptr, b->d_frame_off, b->d_block_clip, total,
                                                             rows, mel, band, n_mels, pre_square, log_power, amin,
                                                             out, clip_max, FC);
    }
    HPSS_LAUNCHED("mask_mel_kernel");
    return HPSS_OK;
}

int launch_topdb(hpss_ctx* ctx, const hpss_batch* b, float* out, int rows_per_stream, int n_streams,
                 const uint32_t* clip_max, float top_db, cudaStream_t st) {
    (void)ctx;
    const int64_t total = b->frame_off[b->n_clips];
    if (total == 0) return HPSS_OK;
    const unsigned grid = (unsigned)((total + 31) / 32);
    topdb_kernel<<<grid, kThreads, 0, st>>>(out, b->d_frame_off, b->d_block_clip, total, rows_per_stream, n_streams,
                                            clip_max, top_db);
    HPSS_LAUNCHED("topdb_kernel");
    return HPSS_OK;
}

}  // namespace hpss
