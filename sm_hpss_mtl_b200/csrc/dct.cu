// MFCC extension: orthonormal DCT-II along the mel axis of a log-mel featuregram,
//   scipy.fftpack.dct(S, axis=0, type=2, norm='ortho')[:n_mfcc]   (what librosa.feature.mfcc applies to
//   power_to_db(melspectrogram)), per stream of the (n_streams * M, T) feature layout.
// The reference itself has no MFCC / DCT: this stage is an extension named by
// BASELINE.json's north_star, pinned to scipy's DCT in the tests ("parity unpinned by the reference").
//
// One warp owns 32 consecutive frames of the batch (lane = frame) and one stream: the M log-mel values of its
// column are read once (coalesced 128-byte row segments), every value feeds n_mfcc accumulators in registers;
// the basis D^T (M x n_mfcc, float32 from a float64 evaluation) sits in shared memory and is read as
// broadcast float4s.  fp32 FMA in increasing m; M x n_mfcc is far too small for tensor-core tiles to pay.
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "common.cuh"

namespace hpss {

namespace {

constexpr int kDctThreads = 128;
constexpr int kDctWarps = kDctThreads / 32;
constexpr int kRows = 8;               // mel rows per load batch
constexpr int kMaxMfcc = 64;           // accumulators per lane (registers)

// NC4: n_mfcc rounded up to a multiple of 4, divided by 4.  COLS: streams per lane (2 = the harmonic and the
// percussive column of one frame share every basis read; halves the shared-memory traffic per FFMA).
template <int NC4, int COLS>
__global__ void __launch_bounds__(kDctThreads)
dct_kernel(const float* __restrict__ feat, const float* __restrict__ basis_t, const int64_t* __restrict__ frame_off,
           const int32_t* __restrict__ block_clip, int64_t total_frames, int M, int n_streams, int n_mfcc,
           float* __restrict__ out) {
    extern __shared__ __align__(16) float s_basis[];       // [M rounded up to kRows][4 * NC4], zero rows at the end
    const int m_pad = (M + kRows - 1) / kRows * kRows;
    for (int i = threadIdx.x; i < m_pad * 4 * NC4; i += kDctThreads) s_basis[i] = (i < M * 4 * NC4) ? basis_t[i] : 0.f;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int groups = n_streams / COLS;                                  // launch guarantees divisibility
    const int64_t n_tasks = ((total_frames + 31) / 32) * groups;          // task = (32-frame block, stream group)
    const float4* b4 = reinterpret_cast<const float4*>(s_basis);
    // the basis is staged once per CTA, so every warp walks many tasks (grid-stride)
    for (int64_t task = (int64_t)blockIdx.x * kDctWarps + warp; task < n_tasks; task += (int64_t)gridDim.x * kDctWarps) {
        const int stream = (int)(task % groups) * COLS;
        const int64_t gf = (task / groups) * 32 + lane;
        if (gf >= total_frames) continue;
        const int c = find_clip_hint(frame_off, block_clip, gf);
        const int64_t fo = __ldg(frame_off + c);
        const int64_t T = __ldg(frame_off + c + 1) - fo;
        const float* x = feat + (int64_t)(n_streams * M) * fo + (int64_t)(stream * M) * T + (gf - fo);
        float* y = out + (int64_t)(n_streams * n_mfcc) * fo + (int64_t)(stream * n_mfcc) * T + (gf - fo);
        const int64_t xs = (int64_t)M * T, ys = (int64_t)n_mfcc * T;      // stream pitch of the input / output
        float acc[COLS][4 * NC4];
#pragma unroll
        for (int s = 0; s < COLS; ++s)
#pragma unroll
            for (int k = 0; k < 4 * NC4; ++k) acc[s][k] = 0.f;
        // rows in batches of kRows; the next batch is requested before the current one is consumed, so that
        // kRows..2*kRows row loads per lane and stream are in flight (the kernel is latency-bound otherwise)
        float cur[COLS][kRows], nxt[COLS][kRows];
#pragma unroll
        for (int s = 0; s < COLS; ++s)
#pragma unroll
            for (int u = 0; u < kRows; ++u) cur[s][u] = (u < M) ? __ldg(x + s * xs + (int64_t)u * T) : 0.f;
#pragma unroll 1
        for (int m0 = 0; m0 < M; m0 += kRows) {
            const float* xn = x + (int64_t)(m0 + kRows) * T;
#pragma unroll
            for (int s = 0; s < COLS; ++s)
#pragma unroll
                for (int u = 0; u < kRows; ++u) nxt[s][u] = (m0 + kRows + u < M) ? __ldg(xn + s * xs + (int64_t)u * T) : 0.f;
#pragma unroll
            for (int u = 0; u < kRows; ++u) {
#pragma unroll
                for (int k = 0; k < NC4; ++k) {
                    const float4 w = b4[(m0 + u) * NC4 + k];       // rows >= M of the staged basis are zero
#pragma unroll
                    for (int s = 0; s < COLS; ++s) {
                        const float v = cur[s][u];
                        acc[s][4 * k + 0] = fmaf(w.x, v, acc[s][4 * k + 0]);
                        acc[s][4 * k + 1] = fmaf(w.y, v, acc[s][4 * k + 1]);
                        acc[s][4 * k + 2] = fmaf(w.z, v, acc[s][4 * k + 2]);
                        acc[s][4 * k + 3] = fmaf(w.w, v, acc[s][4 * k + 3]);
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < COLS; ++s)
#pragma unroll
                for (int u = 0; u < kRows; ++u) cur[s][u] = nxt[s][u];
        }
#pragma unroll
        for (int s = 0; s < COLS; ++s)
#pragma unroll
            for (int k = 0; k < 4 * NC4; ++k)
                if (k < n_mfcc) y[s * ys + (int64_t)k * T] = acc[s][k];
    }
}

}  // namespace

// D^T[m][k] = s_k * cos(pi * k * (2m + 1) / (2M)), s_0 = sqrt(1/M), s_k = sqrt(2/M); padded to a multiple of 4 columns
void build_dct_basis_t(int M, int n_mfcc, int ncols, float* out) {
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < ncols; ++k) {
            double v = 0.0;
            if (k < n_mfcc) {
                const double s = (k == 0) ? sqrt(1.0 / M) : sqrt(2.0 / M);
                v = s * cos(M_PI * (double)k * (2.0 * m + 1.0) / (2.0 * M));
            }
            out[(size_t)m * ncols + k] = (float)v;
        }
}

int launch_dct(hpss_ctx* ctx, const hpss_batch* b, const float* feat, int M, int n_streams, int n_mfcc, float* out,
               cudaStream_t st) {
    if (M < 1 || n_streams < 1 || n_mfcc < 1 || n_mfcc > M || n_mfcc > kMaxMfcc) {
        set_error("dct: need 1 <= n_mfcc <= min(M, %d) (got M=%d n_mfcc=%d)", kMaxMfcc, M, n_mfcc);
        return HPSS_ERR_INVALID;
    }
    const int64_t total = b->frame_off[b->n_clips];
    if (total == 0) return HPSS_OK;
    const int nc4 = (n_mfcc + 3) / 4;
    const int ncols = 4 * nc4;
    float* d_basis = nullptr;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        auto key = std::make_pair(M, n_mfcc);
        auto it = ctx->dct_plans.find(key);
        if (it == ctx->dct_plans.end()) {
            std::vector<float> h((size_t)M * ncols);
            build_dct_basis_t(M, n_mfcc, ncols, h.data());
            HPSS_CUDA(cudaMalloc(&d_basis, sizeof(float) * h.size()));
            HPSS_CUDA(cudaMemcpy(d_basis, h.data(), sizeof(float) * h.size(), cudaMemcpyHostToDevice));
            ctx->dct_plans[key] = d_basis;
        } else {
            d_basis = it->second;
        }
    }
    const size_t smem = sizeof(float) * (size_t)((M + kRows - 1) / kRows * kRows) * ncols;
    if (smem > (size_t)ctx->max_smem_optin) {
        set_error("dct: basis of %zu bytes does not fit shared memory", smem);
        return HPSS_ERR_UNSUPPORTED;
    }
    // two streams per lane while the accumulators fit comfortably in registers (n_mfcc <= 24)
    const int cols = (n_streams % 2 == 0 && nc4 <= 6) ? 2 : 1;
    const int64_t n_tasks = ((total + 31) / 32) * (n_streams / cols);
    const int64_t want = (n_tasks + kDctWarps - 1) / kDctWarps;
    const int mult = knobs().dct_grid_mult;
    const unsigned grid = (unsigned)(mult > 0 ? std::min<int64_t>(want, (int64_t)ctx->sm_count * mult) : want);
#define HPSS_DCT_LAUNCH(NC4)                                                                                       \
    case NC4: {                                                                                                    \
        auto kern = (cols == 2 && NC4 <= 6) ? dct_kernel<NC4, (NC4 <= 6 ? 2 : 1)> : dct_kernel<NC4, 1>;            \
        HPSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
        kern<<<grid, kDctThreads, smem, st>>>(feat, d_basis, b->d_frame_off, b->d_block_clip, total, M, n_streams, \
                                              n_mfcc, out);                                                        \
        break;                                                                                                     \
    }
    switch (nc4) {
        HPSS_DCT_LAUNCH(1) HPSS_DCT_LAUNCH(2) HPSS_DCT_LAUNCH(3) HPSS_DCT_LAUNCH(4) HPSS_DCT_LAUNCH(5) HPSS_DCT_LAUNCH(6)
        HPSS_DCT_LAUNCH(7) HPSS_DCT_LAUNCH(8) HPSS_DCT_LAUNCH(9) HPSS_DCT_LAUNCH(10) HPSS_DCT_LAUNCH(11) HPSS_DCT_LAUNCH(12)
        HPSS_DCT_LAUNCH(13) HPSS_DCT_LAUNCH(14) HPSS_DCT_LAUNCH(15) HPSS_DCT_LAUNCH(16)
        default:
            set_error("dct: n_mfcc=%d unsupported", n_mfcc);
            return HPSS_ERR_UNSUPPORTED;
    }
#undef HPSS_DCT_LAUNCH
    HPSS_LAUNCHED("dct_kernel");
    return HPSS_OK;
}

}  // namespace hpss
