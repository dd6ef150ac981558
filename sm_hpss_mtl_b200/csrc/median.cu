// K2: sliding medians of librosa.decompose.hpss, i.e.
//   scipy.ndimage.median_filter(S, size=(1,k), mode='reflect')   (time axis,  harmonic)
//   scipy.ndimage.median_filter(S, size=(k,1), mode='reflect')   (freq axis,  percussive)
// (lib/preprocessing.py:408,418,430,440 of the reference).  Pure selection on float32:
// results are bit-exact.
//
// Work decomposition: one warp owns 32 independent *lines* (time axis: 32 consecutive
// rows (clip, f) of the batch; frequency axis: 32 consecutive frames (clip, t)) and one
// tile of TT output positions along the filter axis.  The TT+k-1 inputs of every line
// (reflected at the clip borders while loading) sit in shared memory, bank-conflict free
// for "lane = line" access.  Each lane then walks its line in groups of G outputs: it
// pulls the k+G-1 inputs of a group into registers and runs a generated min/max selection
// network (tools/gen_median_networks.py: pruned sort of the k-G+1 values common to the G
// windows + pairwise merge-selection of the rest), ~1.8 k min/max per output instead of a
// full sort.  Outputs overwrite the line in place and are stored coalesced.
// The stage is bound by the ALU pipe (FMNMX / FMNMX3), not by HBM.
#include "common.cuh"
#include "median_networks_gen.cuh"

namespace hpss {

namespace {

constexpr int kWarpsPerCta = 4;

struct LineInfo {
    int64_t base;   // element offset of position 0 of this lane's line
    int n;          // line length along the filter axis (0 = lane idle)
    int estride;    // element stride between positions (time axis: 1, freq axis: T_c)
};

template <bool TIME_AXIS>
__device__ __forceinline__ LineInfo lane_line(const int64_t* __restrict__ frame_off, int n_clips, int rows,
                                              int64_t n_lines, int64_t line) {
    LineInfo li;
    li.base = 0; li.n = 0; li.estride = 1;
    if (line >= n_lines) return li;
    if (TIME_AXIS) {
        const int c = (int)(line / rows);
        const int f = (int)(line - (int64_t)c * rows);
        const int64_t fo = __ldg(frame_off + c);
        const int T = (int)(__ldg(frame_off + c + 1) - fo);
        li.base = (int64_t)rows * fo + (int64_t)f * T;
        li.n = T;
        li.estride = 1;
    } else {
        const int c = find_clip(frame_off, n_clips, line);
        const int64_t fo = __ldg(frame_off + c);
        const int T = (int)(__ldg(frame_off + c + 1) - fo);
        li.base = (int64_t)rows * fo + (line - fo);
        li.n = rows;
        li.estride = T;
    }
    return li;
}

// smem index of (lane, pos): time axis keeps a line contiguous (odd stride), frequency
// axis keeps a position contiguous across lanes.
template <bool TIME_AXIS>
__device__ __forceinline__ int sidx(int lane, int pos, int lstride) {
    return TIME_AXIS ? lane * lstride + pos : pos * 32 + lane;
}

template <bool TIME_AXIS>
__device__ __forceinline__ void tile_load(float* __restrict__ sm, const float* __restrict__ S, const LineInfo& li,
                                          int lane, int p0, int halo, int span, int lstride) {
    if (TIME_AXIS) {
        for (int r = 0; r < 32; ++r) {
            const int64_t b = __shfl_sync(0xffffffffu, li.base, r);
            const int n = __shfl_sync(0xffffffffu, li.n, r);
            if (n == 0 || p0 >= n) continue;
            for (int pos = lane; pos < span; pos += 32) {
                const int src = reflect_idx(p0 - halo + pos, n);
                sm[r * lstride + pos] = __ldg(S + b + src);
            }
        }
    } else {
        if (li.n > 0) {
            for (int pos = 0; pos < span; ++pos) {
                const int src = reflect_idx(p0 - halo + pos, li.n);
                sm[pos * 32 + lane] = __ldg(S + li.base + (int64_t)src * li.estride);
            }
        }
    }
}

template <bool TIME_AXIS>
__device__ __forceinline__ void tile_store(const float* __restrict__ sm, float* __restrict__ out, const LineInfo& li,
                                           int lane, int p0, int TT, int lstride) {
    if (TIME_AXIS) {
        for (int r = 0; r < 32; ++r) {
            const int64_t b = __shfl_sync(0xffffffffu, li.base, r);
            const int n = __shfl_sync(0xffffffffu, li.n, r);
            if (n == 0 || p0 >= n) continue;
            const int lim = min(TT, n - p0);
            for (int pos = lane; pos < lim; pos += 32) out[b + p0 + pos] = sm[r * lstride + pos];
        }
    } else {
        if (li.n > 0) {
            const int lim = min(TT, li.n - p0);
            for (int pos = 0; pos < lim; ++pos)
                out[li.base + (int64_t)(p0 + pos) * li.estride] = sm[pos * 32 + lane];
        }
    }
}

// ---- fast path: odd k with a generated selection network ---------------------------
template <int K, bool TIME_AXIS>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
median_fast_kernel(const float* __restrict__ S, float* __restrict__ out, const int64_t* __restrict__ frame_off,
                   int n_clips, int rows, int64_t n_lines, int TT, int n_ptiles, int64_t n_items) {
    constexpr int G = MedianGroup<K>::G;
    constexpr int HALO = K / 2;
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int span = TT + K - 1;
    const int lstride = span | 1;
    float* sm = smem + (size_t)warp * (TIME_AXIS ? 32 * lstride : span * 32);
    const int NG = TT / G;

    for (int64_t item = (int64_t)blockIdx.x * kWarpsPerCta + warp; item < n_items;
         item += (int64_t)gridDim.x * kWarpsPerCta) {
        const int64_t lb = item / n_ptiles;
        const int p0 = (int)(item - lb * n_ptiles) * TT;
        const LineInfo li = lane_line<TIME_AXIS>(frame_off, n_clips, rows, n_lines, lb * 32 + lane);
        // warp-uniform early exit: tile starts beyond every line of this warp
        const bool live = li.n > 0 && p0 < li.n;
        if (!__any_sync(0xffffffffu, live)) continue;
        tile_load<TIME_AXIS>(sm, S, li, lane, p0, HALO, span, lstride);
        __syncwarp();
        if (live) {
            const int ng = min(NG, (li.n - p0 + G - 1) / G);
            for (int g = 0; g < ng; ++g) {
                float x[K + G - 1], o[G];
#pragma unroll
                for (int i = 0; i < K + G - 1; ++i) x[i] = sm[sidx<TIME_AXIS>(lane, g * G + i, lstride)];
                MedianGroup<K>::run(x, o);
#pragma unroll
                for (int j = 0; j < G; ++j) sm[sidx<TIME_AXIS>(lane, g * G + j, lstride)] = o[j];
            }
        }
        __syncwarp();
        tile_store<TIME_AXIS>(sm, out, li, lane, p0, TT, lstride);
        __syncwarp();
    }
}

// ---- generic path: any k (even k, k > 63): rank counting, O(k^2) per output ---------
template <bool TIME_AXIS>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
median_generic_kernel(const float* __restrict__ S, float* __restrict__ out, const int64_t* __restrict__ frame_off,
                      int n_clips, int rows, int64_t n_lines, int k, int TT, int n_ptiles, int64_t n_items) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int halo = k / 2, rank = k / 2;
    const int span = TT + k - 1;
    const int lstride = span | 1;
    float* sm = smem + (size_t)warp * (TIME_AXIS ? 32 * lstride : span * 32);

    for (int64_t item = (int64_t)blockIdx.x * kWarpsPerCta + warp; item < n_items;
         item += (int64_t)gridDim.x * kWarpsPerCta) {
        const int64_t lb = item / n_ptiles;
        const int p0 = (int)(item - lb * n_ptiles) * TT;
        const LineInfo li = lane_line<TIME_AXIS>(frame_off, n_clips, rows, n_lines, lb * 32 + lane);
        const bool live = li.n > 0 && p0 < li.n;
        if (!__any_sync(0xffffffffu, live)) continue;
        tile_load<TIME_AXIS>(sm, S, li, lane, p0, halo, span, lstride);
        __syncwarp();
        if (live) {
            const int lim = min(TT, li.n - p0);
            for (int p = 0; p < lim; ++p) {
                float res = 0.f;
                for (int j = 0; j < k; ++j) {
                    const float v = sm[sidx<TIME_AXIS>(lane, p + j, lstride)];
                    int lt = 0, le = 0;
                    for (int i = 0; i < k; ++i) {
                        const float w = sm[sidx<TIME_AXIS>(lane, p + i, lstride)];
                        lt += (w < v);
                        le += (w <= v);
                    }
                    if (lt <= rank && rank < le) { res = v; break; }
                }
                // in place: later windows of this line start at p+1
                sm[sidx<TIME_AXIS>(lane, p, lstride)] = res;
            }
        }
        __syncwarp();
        tile_store<TIME_AXIS>(sm, out, li, lane, p0, TT, lstride);
        __syncwarp();
    }
}

template <int K, bool TIME_AXIS>
int launch_fast(hpss_ctx* ctx, const float* S, float* out, const int64_t* d_frame_off, int n_clips, int rows,
                int64_t n_lines, int64_t max_len, cudaStream_t st) {
    constexpr int G = MedianGroup<K>::G;
    const int tt_max = 16 * G;
    const int64_t nt0 = (max_len + tt_max - 1) / tt_max;
    int TT = (int)((max_len + nt0 - 1) / nt0);
    TT = (TT + G - 1) / G * G;
    const int n_ptiles = (int)((max_len + TT - 1) / TT);
    const int64_t n_lb = (n_lines + 31) / 32;
    const int64_t n_items = n_lb * n_ptiles;
    const int span = TT + K - 1;
    const size_t per_warp = (TIME_AXIS ? (size_t)32 * (span | 1) : (size_t)span * 32) * sizeof(float);
    const size_t smem = per_warp * kWarpsPerCta;
    auto kern = median_fast_kernel<K, TIME_AXIS>;
    HPSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    HPSS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kWarpsPerCta * 32, smem));
    if (occ < 1) occ = 1;
    int64_t grid = (n_items + kWarpsPerCta - 1) / kWarpsPerCta;
    const int64_t cap = (int64_t)ctx->sm_count * occ;
    if (grid > cap) grid = cap;
    kern<<<(unsigned)grid, kWarpsPerCta * 32, smem, st>>>(S, out, d_frame_off, n_clips, rows, n_lines, TT,
                                                           n_ptiles, n_items);
    HPSS_LAUNCHED("median_fast_kernel");
    return HPSS_OK;
}

template <bool TIME_AXIS>
int launch_generic(hpss_ctx* ctx, const float* S, float* out, const int64_t* d_frame_off, int n_clips, int rows,
                   int64_t n_lines, int64_t max_len, int k, cudaStream_t st) {
    // per-warp tile of at most ~40 KB of shared memory
    int tt_max = 320 - k;
    if (tt_max < 8) {
        set_error("median kernel size k=%d is too large (max 311)", k);
        return HPSS_ERR_UNSUPPORTED;
    }
    if (tt_max > 128) tt_max = 128;
    const int64_t nt0 = (max_len + tt_max - 1) / tt_max;
    const int TT = (int)((max_len + nt0 - 1) / nt0);
    const int n_ptiles = (int)((max_len + TT - 1) / TT);
    const int64_t n_lb = (n_lines + 31) / 32;
    const int64_t n_items = n_lb * n_ptiles;
    const int span = TT + k - 1;
    const size_t per_warp = (TIME_AXIS ? (size_t)32 * (span | 1) : (size_t)span * 32) * sizeof(float);
    const size_t smem = per_warp * kWarpsPerCta;
    auto kern = median_generic_kernel<TIME_AXIS>;
    HPSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 1;
    HPSS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kWarpsPerCta * 32, smem));
    if (occ < 1) occ = 1;
    int64_t grid = (n_items + kWarpsPerCta - 1) / kWarpsPerCta;
    const int64_t cap = (int64_t)ctx->sm_count * occ;
    if (grid > cap) grid = cap;
    kern<<<(unsigned)grid, kWarpsPerCta * 32, smem, st>>>(S, out, d_frame_off, n_clips, rows, n_lines, k, TT,
                                                           n_ptiles, n_items);
    HPSS_LAUNCHED("median_generic_kernel");
    return HPSS_OK;
}

}  // namespace

int launch_median(hpss_ctx* ctx, const hpss_batch* b, const float* S, int rows, int k, bool time_axis, float* out,
                  cudaStream_t st) {
    if (k < 1 || rows < 1) {
        set_error("median: k=%d rows=%d must be positive", k, rows);
        return HPSS_ERR_INVALID;
    }
    const int64_t total_frames = b->frame_off[b->n_clips];
    if (total_frames == 0) return HPSS_OK;
    const int64_t n_lines = time_axis ? (int64_t)b->n_clips * rows : total_frames;
    const int64_t max_len = time_axis ? b->max_frames : rows;
    if (k == 1) {
        HPSS_CUDA(cudaMemcpyAsync(out, S, sizeof(float) * (size_t)rows * total_frames, cudaMemcpyDeviceToDevice, st));
        return HPSS_OK;
    }
#define HPSS_DISPATCH_K(KK)                                                                                   \
    if (k == KK) {                                                                                            \
        return time_axis ? launch_fast<KK, true>(ctx, S, out, b->d_frame_off, b->n_clips, rows, n_lines,      \
                                                 max_len, st)                                                 \
                         : launch_fast<KK, false>(ctx, S, out, b->d_frame_off, b->n_clips, rows, n_lines,     \
                                                  max_len, st);                                               \
    }
    HPSS_MEDIAN_FAST_KS(HPSS_DISPATCH_K)
#undef HPSS_DISPATCH_K
    return time_axis ? launch_generic<true>(ctx, S, out, b->d_frame_off, b->n_clips, rows, n_lines, max_len, k, st)
                     : launch_generic<false>(ctx, S, out, b->d_frame_off, b->n_clips, rows, n_lines, max_len, k, st);
}

}  // namespace hpss
