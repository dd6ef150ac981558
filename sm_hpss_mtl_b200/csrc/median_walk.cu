// K2p (and K2p + K3 fused): frequency-axis sliding median as a register walk, for the kernel sizes that have a
// stateful double-step selection network (K = 4G - 1: 15, 31; tools/gen_median_networks.py, gen_step).
//
// scipy.ndimage.median_filter(S, size=(k,1), mode='reflect') inside librosa.decompose.hpss
// (lib/preprocessing.py:408,418,430,440), optionally followed in the same kernel by the soft masks, S*mask,
// the Slaney mel projection and power_to_db of K3 (maskmel.cu, mask_mel_sweep_kernel).
//
// One warp owns 32 consecutive frames of the batch (lane = frame) and walks the frequency axis upwards, 2G
// outputs per step.  With lanes along time every global access is a coalesced 128-byte row segment, so there
// is no shared-memory staging, no loader warp and no barrier: the 2G new input rows of the next step are
// prefetched into registers while the selection network of the current step runs.  A step receives the two
// sorted blocks it shares with the previous step, sorts two new blocks and merges (see gen_step); the raw
// values a later step needs again (window edges) are carried in registers.  Reflection at the frequency
// borders is a warp-uniform index computation.
//
// FUSED: each step's 2G percussive medians meet S (already in registers: the window centres) and the harmonic
// median (one more coalesced load per row) in softmask_apply, and the masked values go straight into the mel
// sweep of K3: the percussive spectrogram and both masked spectrograms never exist in memory.  The mask / mel /
// log arithmetic runs on the FMA and MUFU pipes next to the FMNMX stream that saturates the ALU pipe.
#include "maskmath.cuh"
#include "median_networks_gen.cuh"

namespace hpss {

namespace {

#ifndef HPSS_WALK_MINB
#define HPSS_WALK_MINB 4
#endif
constexpr int kWalkWarps = 4;

// p + i * pitch_bytes as ONE IMAD.WIDE.U32 (FMA pipe): the ALU pipe belongs to the FMNMX stream
__device__ __forceinline__ const float* row_ptr(const float* p, uint32_t i, uint32_t pitch_bytes) {
    return reinterpret_cast<const float*>(reinterpret_cast<const char*>(p) + (uint64_t)i * pitch_bytes);
}
__device__ __forceinline__ float* row_ptr(float* p, uint32_t i, uint32_t pitch_bytes) {
    return reinterpret_cast<float*>(reinterpret_cast<char*>(p) + (uint64_t)i * pitch_bytes);
}

struct WalkArgs {
    const float* S;
    float* perc;             // !FUSED: output (rows, T_c) per clip
    const float* harm;       // FUSED
    float* feat;             // FUSED: (2 * n_mels, T_c) per clip
    uint32_t* clip_max;      // FUSED, may be null
    const uint32_t* emit4;   // FUSED: filters finishing before each row, 4 bits per row (MelPlan::d_emit4)
    const float2* sweep_w;   // FUSED: weights of the two open filters per row (MelPlan::d_sweep_w)
    int n_mels;
    int log_power;
    float amin;
};

template <int K, bool FUSED, int LOGP>
__global__ void __launch_bounds__(kWalkWarps * 32, (FUSED || K > 35) ? (K > 47 ? 2 : 3) : HPSS_WALK_MINB)
median_freq_walk_kernel(WalkArgs a, const int64_t* __restrict__ frame_off, const int32_t* __restrict__ block_clip,
                        int64_t total_frames, int rows) {
    using Step = MedianStep<K>;
    constexpr int G = Step::G;
    constexpr int HALO = K / 2;               // = 2G - 1
    constexpr int NR = Step::NRAW;            // = 6G - 11 ... raw inputs of a step, in step_raw_index order
    static_assert(!FUSED || (2 * G) % 8 == 0, "a fused step covers whole words of the emission table");
    static_assert(K == 4 * G - 1 && NR == 2 * (G - 1) + 2 * G + (G - 1), "stateful step layout");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t g0 = ((int64_t)blockIdx.x * kWalkWarps + warp) * 32;
    if (g0 >= total_frames) return;
    const int64_t gf = g0 + lane;
    const bool valid = gf < total_frames;
    int clip = 0;
    int64_t T = 1, fo = 0;
    if (valid) {
        clip = find_clip_hint(frame_off, block_clip, gf);
        fo = __ldg(frame_off + clip);
        T = __ldg(frame_off + clip + 1) - fo;
    }
    const int64_t in_base = (int64_t)rows * fo + (gf - fo);
    const float* col = a.S + in_base;
    const int Ti = (int)T;                     // row pitch of this lane's clip
    const uint32_t T4 = 4u * (uint32_t)Ti;     // ... in bytes: one IMAD.WIDE.U32 per address
    // S[f] of this lane's frame, f reflected into [0, rows) (warp-uniform index)
    auto ld = [&](int f) -> float {
        const int fr = reflect_idx(f, rows);
        return valid ? __ldg(row_ptr(col, (uint32_t)fr, T4)) : 0.f;
    };

    // ---- FUSED state (mel sweep of K3)
    const float* hcol = FUSED ? a.harm + in_base : nullptr;
    float* oh = nullptr;
    float* op = nullptr;
    if (FUSED) {
        oh = a.feat + (int64_t)(2 * a.n_mels) * fo + (gf - fo);
        op = oh + (int64_t)a.n_mels * T;
    }
    float* pcol = FUSED ? nullptr : a.perc + in_base;
    // per-warp scratch of the fused variant: the masked values of one step, [2][2G][32]
    __shared__ float s_scr[FUSED ? kWalkWarps * 4 * G * 32 : 1];
    float* scr = s_scr + (FUSED ? warp * 4 * G * 32 : 0);
    float aH = 0.f, aP = 0.f, bH = 0.f, bP = 0.f;
    float vmaxH = -INFINITY, vmaxP = -INFINITY;
    int cur = 0;
    auto emit = [&]() {
        const float vH = post_value(aH, LOGP, a.amin);
        const float vP = post_value(aP, LOGP, a.amin);
        if (valid) { *oh = vH; *op = vP; }
        oh += T; op += T;
        vmaxH = fmaxf(vmaxH, vH);
        vmaxP = fmaxf(vmaxP, vP);
        aH = bH; aP = bP; bH = 0.f; bP = 0.f;
        ++cur;
    };

    // ---- prologue: x[i] = S[-HALO + i], i = 0 .. 2G + K - 2 of step 0
    // raw values carried between steps: lx = x[0..G-2], mid = x[G..2G-2], c1 = x[2G-1..3G-2], hi = x[3G-1..4G-3]
    float lx[G - 1], mid[G - 1], c1[G], hi[G - 1];
    float ca[G], cb[G];
    {
        float r0[G];
#pragma unroll
        for (int i = 0; i < G - 1; ++i) lx[i] = ld(-HALO + i);
        r0[0] = ld(-HALO + G - 1);
#pragma unroll
        for (int i = 0; i < G - 1; ++i) { mid[i] = ld(-HALO + G + i); r0[1 + i] = mid[i]; }
#pragma unroll
        for (int i = 0; i < G; ++i) c1[i] = ld(-HALO + 2 * G - 1 + i);          // x[2G-1 .. 3G-2] = S[0 .. G-1]
#pragma unroll
        for (int i = 0; i < G - 1; ++i) hi[i] = ld(-HALO + 3 * G - 1 + i);      // x[3G-1 .. 4G-3] = S[G .. 2G-2]
        Step::sort(r0, ca);
        Step::sort(c1, cb);
    }
    float nw[2 * G];                                      // x[4G-2 .. 6G-3] = S[base + 2G-1 .. base + 4G-2]
#pragma unroll
    for (int i = 0; i < 2 * G; ++i) nw[i] = ld(2 * G - 1 + i);

    const int nsteps = (rows + 2 * G - 1) / (2 * G);
#pragma unroll 1
    for (int s = 0; s < nsteps; ++s) {
        const int base = 2 * G * s;
        const bool interior = base + 6 * G - 2 < rows;    // no reflection in this step's loads (warp-uniform)
        // harmonic medians at this step's output rows: in flight while the selection network runs
        float hv[2 * G];
        uint32_t em[(2 * G + 7) / 8];
        if (FUSED) {
#pragma unroll
            for (int i = 0; i < (2 * G + 7) / 8; ++i) em[i] = __ldg(a.emit4 + (base >> 3) + i);
            const float* hp = row_ptr(hcol, (uint32_t)base, T4);
            if (interior) {
#pragma unroll
                for (int j = 0; j < 2 * G; ++j) hv[j] = valid ? __ldg(row_ptr(hp, j, T4)) : 0.f;
            } else {
#pragma unroll
                for (int j = 0; j < 2 * G; ++j) hv[j] = (valid && base + j < rows) ? __ldg(row_ptr(hp, j, T4)) : 0.f;
            }
        }
        float xr[NR], o[2 * G], na[G], nb[G];
#pragma unroll
        for (int i = 0; i < G - 1; ++i) { xr[i] = lx[i]; xr[G - 1 + i] = mid[i]; xr[2 * G - 2 + i] = hi[i]; }
#pragma unroll
        for (int i = 0; i < 2 * G; ++i) xr[3 * G - 3 + i] = nw[i];
        Step::run(ca, cb, xr, o, na, nb);
#pragma unroll
        for (int i = 0; i < G; ++i) { ca[i] = na[i]; cb[i] = nb[i]; }
        // the 2G new input rows of the next step: in flight during the stores / the mask and mel phase
        float nn[2 * G];
        if (interior) {
            const float* np = row_ptr(col, (uint32_t)(base + 4 * G - 1), T4);
#pragma unroll
            for (int i = 0; i < 2 * G; ++i) nn[i] = valid ? __ldg(row_ptr(np, i, T4)) : 0.f;
        } else if (s + 1 < nsteps) {
#pragma unroll
            for (int i = 0; i < 2 * G; ++i) nn[i] = ld(base + 4 * G - 1 + i);
        }

        if (!FUSED) {
            if (valid) {
                float* dst = row_ptr(pcol, (uint32_t)base, T4);
                if (interior) {
#pragma unroll
                    for (int j = 0; j < 2 * G; ++j) *row_ptr(dst, j, T4) = o[j];
                } else {
#pragma unroll
                    for (int j = 0; j < 2 * G; ++j)
                        if (base + j < rows) *row_ptr(dst, j, T4) = o[j];
                }
            }
        } else {
            // S at the output rows = the window centres x[HALO + j] = c1[0..G-1], hi[0..G-2], nw[0]
            {
                float sc[2 * G], Hm[2 * G], Pm[2 * G];
#pragma unroll
                for (int j = 0; j < G; ++j) sc[j] = c1[j];
#pragma unroll
                for (int j = 0; j < G - 1; ++j) sc[G + j] = hi[j];
                sc[2 * G - 1] = nw[0];
                softmask_batch<2 * G>(sc, hv, o, Hm, Pm);
#pragma unroll
                for (int j = 0; j < 2 * G; ++j) {
                    scr[j * 32 + lane] = Hm[j];
                    scr[(2 * G + j) * 32 + lane] = Pm[j];
                }
            }
            __syncwarp();
            // mel sweep, one row per iteration of a rolled loop (its body exists once in the code): the filters
            // that finish before the row (count from the emission table: no load sits in front of a branch),
            // then four FMAs; the next row's weights and masked values are fetched one iteration ahead
            float2 wn = __ldg(a.sweep_w + base);                                    // table is zero padded
            float hnx = scr[lane], pnx = scr[2 * G * 32 + lane];
#pragma unroll 1
            for (int j = 0; j < 2 * G; ++j) {
                const float2 w = wn;
                const float H = hnx, P = pnx;
                wn = __ldg(a.sweep_w + base + j + 1);
                hnx = scr[((j + 1) & (2 * G - 1)) * 32 + lane];
                pnx = scr[(2 * G + ((j + 1) & (2 * G - 1))) * 32 + lane];
                int n = (int)((em[(2 * G > 8 && j >= 8) ? 1 : 0] >> (4 * (j & 7))) & 15u);    // warp-uniform
#pragma unroll 1
                for (; n > 0; --n) emit();
                aH = fmaf(w.x, H, aH);
                aP = fmaf(w.x, P, aP);
                bH = fmaf(w.y, H, bH);
                bP = fmaf(w.y, P, bP);
            }
            __syncwarp();
        }

        // carry the raw values the next step reads again: x'[i] = x[i + 2G]
#pragma unroll
        for (int i = 0; i < G - 1; ++i) lx[i] = c1[1 + i];                      // x[2G .. 3G-2]
#pragma unroll
        for (int i = 0; i < G - 2; ++i) mid[i] = hi[1 + i];                     // x[3G .. 4G-3]
        mid[G - 2] = nw[0];                                                     // x[4G-2]
#pragma unroll
        for (int i = 0; i < G; ++i) c1[i] = nw[1 + i];                          // x[4G-1 .. 5G-2]
#pragma unroll
        for (int i = 0; i < G - 1; ++i) hi[i] = nw[G + 1 + i];                  // x[5G-1 .. 6G-3]
#pragma unroll
        for (int i = 0; i < 2 * G; ++i) nw[i] = nn[i];
    }

    if (FUSED) {
#pragma unroll 1
        while (cur < a.n_mels) emit();                    // filters above the last frequency row
        if (a.clip_max != nullptr) {
            publish_max(a.clip_max, 2, 0, valid, clip, vmaxH);
            publish_max(a.clip_max, 2, 1, valid, clip, vmaxP);
        }
    }
}

// The same register walk for every other kernel size with a generated (stateless) group network: the K + G - 1
// inputs of a group live in registers, a group produces G outputs, the window then moves up by G rows (K - 1
// register moves, G coalesced loads prefetched during the previous network).
template <int K>
__global__ void __launch_bounds__(kWalkWarps * 32, K > 43 ? 3 : 4)
median_freq_walk_group_kernel(const float* __restrict__ S, float* __restrict__ perc, const int64_t* __restrict__ frame_off,
                              const int32_t* __restrict__ block_clip, int64_t total_frames, int rows) {
    constexpr int G = MedianGroup<K>::G;
    constexpr int HALO = K / 2;
    constexpr int NX = K + G - 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t g0 = ((int64_t)blockIdx.x * kWalkWarps + warp) * 32;
    if (g0 >= total_frames) return;
    const int64_t gf = g0 + lane;
    const bool valid = gf < total_frames;
    int64_t fo = 0;
    int Ti = 1;
    if (valid) {
        const int clip = find_clip_hint(frame_off, block_clip, gf);
        fo = __ldg(frame_off + clip);
        Ti = (int)(__ldg(frame_off + clip + 1) - fo);
    }
    const int64_t in_base = (int64_t)rows * fo + (gf - fo);
    const float* col = S + in_base;
    float* pcol = perc + in_base;
    const uint32_t T4 = 4u * (uint32_t)Ti;
    auto ld = [&](int f) -> float {
        const int fr = reflect_idx(f, rows);
        return valid ? __ldg(row_ptr(col, (uint32_t)fr, T4)) : 0.f;
    };
    float x[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = ld(-HALO + i);
    const int ngroups = (rows + G - 1) / G;
#pragma unroll 1
    for (int g = 0; g < ngroups; ++g) {
        const int base = g * G;
        const int fnew = base + HALO + G;                   // first new row of the next group
        const bool interior = fnew + G - 1 < rows;          // warp-uniform: no reflection in the next loads
        float nn[G];
        if (interior) {
            const float* np = row_ptr(col, (uint32_t)fnew, T4);
#pragma unroll
            for (int j = 0; j < G; ++j) nn[j] = valid ? __ldg(row_ptr(np, j, T4)) : 0.f;
        } else if (g + 1 < ngroups) {
#pragma unroll
            for (int j = 0; j < G; ++j) nn[j] = ld(fnew + j);
        }
        float o[G];
        MedianGroup<K>::run(x, o);
        if (valid) {
            float* dst = row_ptr(pcol, (uint32_t)base, T4);
#pragma unroll
            for (int j = 0; j < G; ++j)
                if (base + j < rows) *row_ptr(dst, j, T4) = o[j];
        }
#pragma unroll
        for (int i = 0; i < K - 1; ++i) x[i] = x[i + G];
#pragma unroll
        for (int j = 0; j < G; ++j) x[K - 1 + j] = nn[j];
    }
}

template <int K, bool FUSED>
int launch_walk(const hpss_batch* b, const WalkArgs& wa, int rows, int64_t total, cudaStream_t st) {
    const int64_t n_warps = (total + 31) / 32;
    const unsigned grid = (unsigned)((n_warps + kWalkWarps - 1) / kWalkWarps);
    if (FUSED && wa.log_power)
        median_freq_walk_kernel<K, FUSED, 1><<<grid, kWalkWarps * 32, 0, st>>>(wa, b->d_frame_off, b->d_block_clip, total, rows);
    else
        median_freq_walk_kernel<K, FUSED, 0><<<grid, kWalkWarps * 32, 0, st>>>(wa, b->d_frame_off, b->d_block_clip, total, rows);
    HPSS_LAUNCHED("median_freq_walk_kernel");
    return HPSS_OK;
}

}  // namespace

// frequency-axis median as a register walk; *handled = false when k has no stateful step network
int launch_median_freq_walk(hpss_ctx* ctx, const hpss_batch* b, const float* S, int rows, int k, float* out,
                            cudaStream_t st, bool* handled) {
    (void)ctx;
    *handled = false;
    const int64_t total = b->frame_off[b->n_clips];
    WalkArgs wa{};
    wa.S = S; wa.perc = out;
#define HPSS_WALK_ANY_K(KK)                                                                                         \
    if (k == KK) {                                                                                                  \
        *handled = true;                                                                                            \
        if (total == 0) return HPSS_OK;                                                                             \
        if constexpr (MedianStep<KK>::available) {                                                                  \
            return launch_walk<KK, false>(b, wa, rows, total, st);                                                  \
        } else {                                                                                                    \
            const int64_t n_warps = (total + 31) / 32;                                                              \
            const unsigned grid = (unsigned)((n_warps + kWalkWarps - 1) / kWalkWarps);                              \
            median_freq_walk_group_kernel<KK><<<grid, kWalkWarps * 32, 0, st>>>(S, out, b->d_frame_off,             \
                                                                                b->d_block_clip, total, rows);      \
            HPSS_LAUNCHED("median_freq_walk_group_kernel");                                                         \
            return HPSS_OK;                                                                                         \
        }                                                                                                           \
    }
    HPSS_MEDIAN_FAST_KS(HPSS_WALK_ANY_K)
#undef HPSS_WALK_ANY_K
    return HPSS_OK;
}

// K2p + K3 in one kernel (register walk + mel sweep); *handled = false when k has no stateful step network or
// the mel basis cannot be swept
int launch_perc_mask_mel_walk(hpss_ctx* ctx, const hpss_batch* b, const float* S, const float* harm, int rows, int k,
                              const MelPlan* mel, int log_power, float amin, float* out, uint32_t* clip_max,
                              cudaStream_t st, bool* handled) {
    (void)ctx;
    *handled = false;
    if (!mel || !mel->sweepable || !mel->walkable) return HPSS_OK;
    if (log_power != 0 && log_power != 1) return HPSS_OK;
    const int64_t total = b->frame_off[b->n_clips];
    WalkArgs wa{};
    wa.S = S; wa.harm = harm; wa.feat = out; wa.clip_max = clip_max; wa.emit4 = mel->d_emit4; wa.sweep_w = mel->d_sweep_w;
    wa.n_mels = mel->n_mels; wa.log_power = log_power; wa.amin = amin;
#define HPSS_WALK_K(KK)                                                              \
    if (k == KK) {                                                                   \
        if constexpr (MedianStep<KK>::available) {                                   \
            *handled = true;                                                         \
            if (clip_max) HPSS_CUDA(cudaMemsetAsync(clip_max, 0, sizeof(uint32_t) * 2 * (size_t)b->n_clips, st)); \
            if (total == 0) return HPSS_OK;                                          \
            return launch_walk<KK, true>(b, wa, rows, total, st);                    \
        }                                                                            \
    }
    HPSS_WALK_K(15) HPSS_WALK_K(31)
#undef HPSS_WALK_K
    return HPSS_OK;
}

}  // namespace hpss
