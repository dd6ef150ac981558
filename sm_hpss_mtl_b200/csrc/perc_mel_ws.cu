// K2p + K3 fused, warp specialised: the frequency-axis median of librosa.decompose.hpss and everything behind it
// (soft masks, S*mask, Slaney mel sweep, power_to_db without the clip) in one kernel, so that the percussive
// median and both masked spectrograms never exist in memory (lib/preprocessing.py:418-422 of the reference).
//
// A CTA holds 4 producer warps and 4 consumer warps; producer w and consumer w share 32 consecutive frames of
// the batch (lane = frame).  The producer is the register walk of median_walk.cu (stateful double steps, 2G
// outputs per step): it parks the 2G medians of a step and S at the same rows in a double-buffered shared-memory
// slot and signals a `full` mbarrier.  The consumer meanwhile has the harmonic medians of those rows in flight,
// waits, evaluates the soft masks (softmask_batch, bit-identical to numpy) and feeds the mel sweep of
// mask_mel_sweep2_kernel, then releases the slot through an `empty` mbarrier.  The two instruction streams are
// complementary -- FMNMX on the half-rate ALU pipe against FMA / MUFU / load-store work -- and, being separate
// warps, they overlap at instruction granularity, which the single-warp fusion (median_freq_walk_kernel<K,true>)
// cannot do: there the serial sweep phase and the selection network of one warp alternate.
// Results are bit-identical to hpss_median_freq followed by hpss_mask_mel_log_sr (tested).
#include "maskmath.cuh"
#include "median_networks_gen.cuh"

namespace hpss {

namespace {

constexpr int kPairs = 4;                       // producer / consumer warp pairs per CTA
constexpr int kWsThreads = 2 * kPairs * 32;

__device__ __forceinline__ uint32_t ws_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ws_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void ws_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void ws_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WS_WAIT:\n"
        "mbarrier.try_wait.parity.shared.b64 p, [%0], %1;\n"
        "@p bra WS_DONE;\n"
        "bra WS_WAIT;\n"
        "WS_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ const float* ws_row_ptr(const float* p, uint32_t i, uint32_t pitch_bytes) {
    return reinterpret_cast<const float*>(reinterpret_cast<const char*>(p) + (uint64_t)i * pitch_bytes);
}

struct WsArgs {
    const float* S;
    const float* harm;
    float* feat;             // (2 * n_mels, T_c) per clip
    uint32_t* clip_max;      // may be null
    const uint32_t* emit4;   // MelPlan::d_emit4
    const float2* sweep_w;   // MelPlan::d_sweep_w
    int n_mels;
    float amin;
};

template <int K, int LOGP>
__global__ void __launch_bounds__(kWsThreads, 2)
perc_mel_ws_kernel(WsArgs a, const int64_t* __restrict__ frame_off, const int32_t* __restrict__ block_clip,
                   int64_t total_frames, int rows) {
    using Step = MedianStep<K>;
    constexpr int G = Step::G;
    constexpr int HALO = K / 2;               // = 2G - 1
    constexpr int NR = Step::NRAW;
    constexpr int NO = 2 * G;                 // outputs per step
    static_assert(K == 4 * G - 1 && NO % 4 == 0, "stateful step with a multiple of four outputs");
    // slot: [2 buffers][2 arrays: S centre, perc][NO rows][32 lanes]
    __shared__ float s_slot[kPairs][2][2][NO][32];
    __shared__ __align__(8) uint64_t s_bar[kPairs][4];          // full[2], empty[2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = warp & (kPairs - 1);
    const bool producer = warp < kPairs;
    if (threadIdx.x < kPairs) {
        for (int i = 0; i < 4; ++i) ws_mbar_init(ws_smem_u32(&s_bar[threadIdx.x][i]), 32);
    }
    __syncthreads();
    const uint32_t full0 = ws_smem_u32(&s_bar[pair][0]), empty0 = ws_smem_u32(&s_bar[pair][2]);

    const int64_t g0 = ((int64_t)blockIdx.x * kPairs + pair) * 32;
    if (g0 >= total_frames) return;           // both warps of the pair leave together
    const int64_t gf = g0 + lane;
    const bool valid = gf < total_frames;
    int clip = 0;
    int64_t fo = 0;
    int Ti = 1;
    if (valid) {
        clip = find_clip_hint(frame_off, block_clip, gf);
        fo = __ldg(frame_off + clip);
        Ti = (int)(__ldg(frame_off + clip + 1) - fo);
    }
    const uint32_t T4 = 4u * (uint32_t)Ti;
    const int64_t in_base = (int64_t)rows * fo + (gf - fo);
    const int nsteps = (rows + NO - 1) / NO;

    if (producer) {
        // ===== frequency-axis median walk (see median_freq_walk_kernel) =====
        const float* col = a.S + in_base;
        auto ld = [&](int f) -> float {
            const int fr = reflect_idx(f, rows);
            return valid ? __ldg(ws_row_ptr(col, (uint32_t)fr, T4)) : 0.f;
        };
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
            for (int i = 0; i < G; ++i) c1[i] = ld(-HALO + 2 * G - 1 + i);
#pragma unroll
            for (int i = 0; i < G - 1; ++i) hi[i] = ld(-HALO + 3 * G - 1 + i);
            Step::sort(r0, ca);
            Step::sort(c1, cb);
        }
        float nw[NO];
#pragma unroll
        for (int i = 0; i < NO; ++i) nw[i] = ld(2 * G - 1 + i);
#pragma unroll 1
        for (int s = 0; s < nsteps; ++s) {
            const int base = NO * s;
            const bool interior = base + 6 * G - 2 < rows;
            float xr[NR], o[NO], na[G], nb[G];
#pragma unroll
            for (int i = 0; i < G - 1; ++i) { xr[i] = lx[i]; xr[G - 1 + i] = mid[i]; xr[2 * G - 2 + i] = hi[i]; }
#pragma unroll
            for (int i = 0; i < NO; ++i) xr[3 * G - 3 + i] = nw[i];
            Step::run(ca, cb, xr, o, na, nb);
#pragma unroll
            for (int i = 0; i < G; ++i) { ca[i] = na[i]; cb[i] = nb[i]; }
            float nn[NO];
            if (interior) {
                const float* np = ws_row_ptr(col, (uint32_t)(base + 4 * G - 1), T4);
#pragma unroll
                for (int i = 0; i < NO; ++i) nn[i] = valid ? __ldg(ws_row_ptr(np, i, T4)) : 0.f;
            } else if (s + 1 < nsteps) {
#pragma unroll
                for (int i = 0; i < NO; ++i) nn[i] = ld(base + 4 * G - 1 + i);
            }
            // hand the step over: S at the output rows (window centres) and the medians
            const int buf = s & 1;
            if (s >= 2) ws_mbar_wait(empty0 + 8u * buf, ((s >> 1) - 1) & 1u);
            float (*slot)[NO][32] = s_slot[pair][buf];
#pragma unroll
            for (int j = 0; j < G; ++j) slot[0][j][lane] = c1[j];
#pragma unroll
            for (int j = 0; j < G - 1; ++j) slot[0][G + j][lane] = hi[j];
            slot[0][NO - 1][lane] = nw[0];
#pragma unroll
            for (int j = 0; j < NO; ++j) slot[1][j][lane] = o[j];
            ws_mbar_arrive(full0 + 8u * buf);
            // carry the raw values the next step reads again: x'[i] = x[i + 2G]
#pragma unroll
            for (int i = 0; i < G - 1; ++i) lx[i] = c1[1 + i];
#pragma unroll
            for (int i = 0; i < G - 2; ++i) mid[i] = hi[1 + i];
            mid[G - 2] = nw[0];
#pragma unroll
            for (int i = 0; i < G; ++i) c1[i] = nw[1 + i];
#pragma unroll
            for (int i = 0; i < G - 1; ++i) hi[i] = nw[G + 1 + i];
#pragma unroll
            for (int i = 0; i < NO; ++i) nw[i] = nn[i];
        }
    } else {
        // ===== soft masks + mel sweep + power_to_db (see mask_mel_sweep2_kernel) =====
        const float* hcol = a.harm + in_base;
        float* oh = a.feat + (int64_t)(2 * a.n_mels) * fo + (gf - fo);
        float* op = oh + (int64_t)a.n_mels * Ti;
        float aH = 0.f, aP = 0.f, bH = 0.f, bP = 0.f;
        float vmaxH = -INFINITY, vmaxP = -INFINITY;
        int cur = 0;
        auto emit = [&]() {
            const float vH = post_value(aH, LOGP, a.amin);
            const float vP = post_value(aP, LOGP, a.amin);
            if (valid) { *oh = vH; *op = vP; }
            oh += Ti; op += Ti;
            vmaxH = fmaxf(vmaxH, vH);
            vmaxP = fmaxf(vmaxP, vP);
            aH = bH; aP = bP; bH = 0.f; bP = 0.f;
            ++cur;
        };
        // harmonic medians of step 0
        float hv[NO];
#pragma unroll
        for (int j = 0; j < NO; ++j) hv[j] = (valid && j < rows) ? __ldg(ws_row_ptr(hcol, j, T4)) : 0.f;
#pragma unroll 1
        for (int s = 0; s < nsteps; ++s) {
            const int base = NO * s;
            const int buf = s & 1;
            // next step's harmonic medians: in flight while this step is evaluated
            float hn[NO];
#pragma unroll
            for (int j = 0; j < NO; ++j) {
                const int f = base + NO + j;
                hn[j] = (valid && f < rows) ? __ldg(ws_row_ptr(hcol, (uint32_t)f, T4)) : 0.f;
            }
            ws_mbar_wait(full0 + 8u * buf, (s >> 1) & 1u);
            const float (*slot)[NO][32] = s_slot[pair][buf];
#pragma unroll
            for (int q = 0; q < NO / 4; ++q) {
                float sv[4], pv[4], h4[4], H[4], P[4];
                float2 w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    sv[u] = slot[0][4 * q + u][lane];
                    pv[u] = slot[1][4 * q + u][lane];
                    h4[u] = hv[4 * q + u];
                    w[u] = __ldg(a.sweep_w + base + 4 * q + u);          // zero padded table
                }
                const int f0 = base + 4 * q;
                const uint32_t em = __ldg(a.emit4 + (f0 >> 3)) >> (4 * (f0 & 7));
                softmask_batch<4>(sv, h4, pv, H, P);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    int n = (int)((em >> (4 * u)) & 15u);                 // warp-uniform
#pragma unroll 1
                    for (; n > 0; --n) emit();
                    aH = fmaf(w[u].x, H[u], aH);
                    aP = fmaf(w[u].x, P[u], aP);
                    bH = fmaf(w[u].y, H[u], bH);
                    bP = fmaf(w[u].y, P[u], bP);
                }
            }
            ws_mbar_arrive(empty0 + 8u * buf);                            // every lane has read its slot values
#pragma unroll
            for (int j = 0; j < NO; ++j) hv[j] = hn[j];
        }
#pragma unroll 1
        while (cur < a.n_mels) emit();
        if (a.clip_max != nullptr) {
            publish_max(a.clip_max, 2, 0, valid, clip, vmaxH);
            publish_max(a.clip_max, 2, 1, valid, clip, vmaxP);
        }
    }
}

template <int K>
int launch_ws(const hpss_batch* b, const WsArgs& wa, int log_power, int rows, int64_t total, cudaStream_t st) {
    const int64_t n_warps = (total + 31) / 32;
    const unsigned grid = (unsigned)((n_warps + kPairs - 1) / kPairs);
    if (log_power)
        perc_mel_ws_kernel<K, 1><<<grid, kWsThreads, 0, st>>>(wa, b->d_frame_off, b->d_block_clip, total, rows);
    else
        perc_mel_ws_kernel<K, 0><<<grid, kWsThreads, 0, st>>>(wa, b->d_frame_off, b->d_block_clip, total, rows);
    HPSS_LAUNCHED("perc_mel_ws_kernel");
    return HPSS_OK;
}

}  // namespace

// *handled = false when k has no suitable stateful step network or the mel basis cannot be swept
int launch_perc_mask_mel_ws(hpss_ctx* ctx, const hpss_batch* b, const float* S, const float* harm, int rows, int k,
                            const MelPlan* mel, int log_power, float amin, float* out, uint32_t* clip_max,
                            cudaStream_t st, bool* handled) {
    (void)ctx;
    *handled = false;
    if (!mel || !mel->sweepable || !mel->walkable) return HPSS_OK;
    if (log_power != 0 && log_power != 1) return HPSS_OK;
    const int64_t total = b->frame_off[b->n_clips];
    WsArgs wa{};
    wa.S = S; wa.harm = harm; wa.feat = out; wa.clip_max = clip_max; wa.emit4 = mel->d_emit4; wa.sweep_w = mel->d_sweep_w;
    wa.n_mels = mel->n_mels; wa.amin = amin;
#define HPSS_WS_K(KK)                                                                                               \
    if (k == KK) {                                                                                                  \
        if constexpr (MedianStep<KK>::available) {                                                                  \
            *handled = true;                                                                                        \
            if (clip_max) HPSS_CUDA(cudaMemsetAsync(clip_max, 0, sizeof(uint32_t) * 2 * (size_t)b->n_clips, st));   \
            if (total == 0) return HPSS_OK;                                                                         \
            return launch_ws<KK>(b, wa, log_power, rows, total, st);                                                \
        }                                                                                                           \
    }
    HPSS_WS_K(15) HPSS_WS_K(23) HPSS_WS_K(31)
#undef HPSS_WS_K
    return HPSS_OK;
}

}  // namespace hpss
