// K2p: frequency-axis sliding median as a register walk: the stateful double-step selection network where the
// kernel size has one (tools/gen_median_networks.py, gen_step), the stateless group network otherwise.
//
// scipy.ndimage.median_filter(S, size=(k,1), mode='reflect') inside librosa.decompose.hpss
// (lib/preprocessing.py:408,418,430,440).
//
// One warp owns 32 consecutive frames of the batch (lane = frame) and walks the frequency axis upwards, 2G
// outputs per step.  With lanes along time every global access is a coalesced 128-byte row segment, so there
// is no shared-memory staging, no loader warp and no barrier: the 2G new input rows of the next step are
// prefetched into registers while the selection network of the current step runs.  A step receives the two
// sorted blocks it shares with the previous step, sorts two new blocks and merges (see gen_step); the raw
// values a later step needs again (window edges) are carried in registers.  Reflection at the frequency
// borders is a warp-uniform index computation.
//
// (Round 1 carried three variants that fused K3's masks and mel sweep into this walk; all were bit-identical and
// all slower than the two separate kernels -- 168 registers, 12 warps per SM, the serial sweep does not overlap the
// FMNMX phase -- and were removed; DESIGN.md keeps the measurements.)
#include "common.cuh"
#include "median_networks_gen.cuh"

namespace hpss {

namespace {

#ifndef HPSS_WALK_MINB
#define HPSS_WALK_MINB 4
#endif
#ifndef HPSS_WALKG_MINB
#define HPSS_WALKG_MINB 4
#endif
constexpr int kWalkWarps = 4;

// Addresses come from the FMA pipe (fma_pipe_* in common.cuh): an element of the lane's clip is clip_base[i0 + row * T]
// with a 32-bit index.  The launcher checks rows * max_frames < 2^32.
__device__ __forceinline__ uint32_t row_idx(uint32_t i0, uint32_t row, uint32_t T) { return fma_pipe_mad(T, row, i0); }
__device__ __forceinline__ float load_elem(const float* base, uint32_t idx, uint64_t four) {
    return fma_pipe_load(base, idx, four);
}
__device__ __forceinline__ void store_elem(float* base, uint32_t idx, uint64_t four, float v) {
    fma_pipe_store(base, idx, four, v);
}

// resident CTAs per SM the register allocation aims at: as many as fit without spilling (measured on the 4096 x 1 s
// batch: k = 7 0.155 -> 0.135 ms at 12, k = 15 0.148 -> 0.138 ms at 8, k = 19 0.162 -> 0.155 ms at 6; one step more
// spills and costs 30-90 %; k = 11 needs 60 registers either way)
constexpr int walk_min_blocks(int K) {
    return K > 47 ? 2 : K > 35 ? 3 : K <= 7 ? 12 : K <= 15 ? 8 : K <= 19 ? 6 : HPSS_WALK_MINB;
}

template <int K>
__global__ void __launch_bounds__(kWalkWarps * 32, walk_min_blocks(K))
median_freq_walk_kernel(const float* __restrict__ S, float* __restrict__ perc, const int64_t* __restrict__ frame_off,
                        const int32_t* __restrict__ block_clip, int64_t total_frames, int rows) {
    using Step = MedianStep<K>;
    constexpr int G = Step::G;
    constexpr int HALO = K / 2;               // = 2G - 1
    constexpr int NR = Step::NRAW;            // raw inputs of a step, in step_raw_index order
    static_assert(K == 4 * G - 1 && NR == 2 * (G - 1) + 2 * G + (G - 1), "stateful step layout");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t g0 = ((int64_t)blockIdx.x * kWalkWarps + warp) * 32;
    if (g0 >= total_frames) return;
    // lanes past the last frame of the batch work on the last frame as well (same loads, same values stored to
    // the same addresses): no per-access predicate, whose compares would sit on the ALU pipe
    const int64_t gf = min(g0 + lane, total_frames - 1);
    const int clip = find_clip_hint(frame_off, block_clip, gf);
    const int64_t fo = __ldg(frame_off + clip);
    const int64_t T64 = __ldg(frame_off + clip + 1) - fo;
    const uint32_t T = (uint32_t)T64;                                     // row pitch of this lane's clip
    const uint32_t i0 = (uint32_t)(gf - fo);                              // the lane's frame inside its clip
    const uint64_t four = 4u + ((uint64_t)T64 >> 31);                                 // = 4, in a per-lane register: load_elem()
    const float* __restrict__ Sc = S + (int64_t)rows * fo;
    float* __restrict__ Pc = perc + (int64_t)rows * fo;
    // S[f] of this lane's frame, f reflected into [0, rows) (warp-uniform index)
    auto ld = [&](int f) -> float { return load_elem(Sc, row_idx(i0, (uint32_t)reflect_idx(f, rows), T), four); };

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
        float xr[NR], o[2 * G], na[G], nb[G];
#pragma unroll
        for (int i = 0; i < G - 1; ++i) { xr[i] = lx[i]; xr[G - 1 + i] = mid[i]; xr[2 * G - 2 + i] = hi[i]; }
#pragma unroll
        for (int i = 0; i < 2 * G; ++i) xr[3 * G - 3 + i] = nw[i];
        Step::run(ca, cb, xr, o, na, nb);
#pragma unroll
        for (int i = 0; i < G; ++i) { ca[i] = na[i]; cb[i] = nb[i]; }
        // the 2G new input rows of the next step: in flight during the stores
        float nn[2 * G];
        if (interior) {
            const uint32_t in = row_idx(i0, (uint32_t)(base + 4 * G - 1), T);
#pragma unroll
            for (int i = 0; i < 2 * G; ++i) nn[i] = load_elem(Sc, row_idx(in, i, T), four);
        } else if (s + 1 < nsteps) {
#pragma unroll
            for (int i = 0; i < 2 * G; ++i) nn[i] = ld(base + 4 * G - 1 + i);
        }
        const uint32_t io = row_idx(i0, (uint32_t)base, T);
        if (interior) {
#pragma unroll
            for (int j = 0; j < 2 * G; ++j) store_elem(Pc, row_idx(io, j, T), four, o[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 2 * G; ++j)
                if (base + j < rows) store_elem(Pc, row_idx(io, j, T), four, o[j]);
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
}

// The same register walk for every other kernel size with a generated (stateless) group network: the K + G - 1
// inputs of a group live in registers, a group produces G outputs, the window then moves up by G rows (K - 1
// register moves, G coalesced loads prefetched during the previous network).
template <int K>
__global__ void __launch_bounds__(kWalkWarps * 32, K > 43 ? 3 : (K <= 17 ? 8 : HPSS_WALKG_MINB))   // (k = 17: 0.172 -> 0.160 ms at 8 CTAs per SM; 21 .. 29 do not move)
median_freq_walk_group_kernel(const float* __restrict__ S, float* __restrict__ perc, const int64_t* __restrict__ frame_off,
                              const int32_t* __restrict__ block_clip, int64_t total_frames, int rows) {
    constexpr int G = MedianGroup<K>::G;
    constexpr int HALO = K / 2;
    constexpr int NX = K + G - 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t g0 = ((int64_t)blockIdx.x * kWalkWarps + warp) * 32;
    if (g0 >= total_frames) return;
    const int64_t gf = min(g0 + lane, total_frames - 1);      // see median_freq_walk_kernel
    const int clip = find_clip_hint(frame_off, block_clip, gf);
    const int64_t fo = __ldg(frame_off + clip);
    const int64_t T64 = __ldg(frame_off + clip + 1) - fo;
    const uint32_t T = (uint32_t)T64;
    const uint32_t i0 = (uint32_t)(gf - fo);
    const uint64_t four = 4u + ((uint64_t)T64 >> 31);
    const float* __restrict__ Sc = S + (int64_t)rows * fo;
    float* __restrict__ Pc = perc + (int64_t)rows * fo;
    auto ld = [&](int f) -> float { return load_elem(Sc, row_idx(i0, (uint32_t)reflect_idx(f, rows), T), four); };
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
            const uint32_t in = row_idx(i0, (uint32_t)fnew, T);
#pragma unroll
            for (int j = 0; j < G; ++j) nn[j] = load_elem(Sc, row_idx(in, j, T), four);
        } else if (g + 1 < ngroups) {
#pragma unroll
            for (int j = 0; j < G; ++j) nn[j] = ld(fnew + j);
        }
        float o[G];
        MedianGroup<K>::run(x, o);
        const uint32_t io = row_idx(i0, (uint32_t)base, T);
#pragma unroll
        for (int j = 0; j < G; ++j)
            if (base + j < rows) store_elem(Pc, row_idx(io, j, T), four, o[j]);
#pragma unroll
        for (int i = 0; i < K - 1; ++i) x[i] = x[i + G];
#pragma unroll
        for (int j = 0; j < G; ++j) x[K - 1 + j] = nn[j];
    }
}

}  // namespace

// frequency-axis median as a register walk; *handled = false when k has no stateful step network
int launch_median_freq_walk(hpss_ctx* ctx, const hpss_batch* b, const float* S, int rows, int k, float* out,
                            cudaStream_t st, bool* handled) {
    (void)ctx;
    *handled = false;
    const int64_t total = b->frame_off[b->n_clips];
    if ((int64_t)rows * b->max_frames > 0xffffffffLL) return HPSS_OK;      // 32-bit indices inside a clip (row_idx)
#define HPSS_WALK_ANY_K(KK)                                                                                         \
    if (k == KK) {                                                                                                  \
        *handled = true;                                                                                            \
        if (total == 0) return HPSS_OK;                                                                             \
        const int64_t n_warps = (total + 31) / 32;                                                                  \
        const unsigned grid = (unsigned)((n_warps + kWalkWarps - 1) / kWalkWarps);                                  \
        if constexpr (MedianStep<KK>::available) {                                                                  \
            median_freq_walk_kernel<KK><<<grid, kWalkWarps * 32, 0, st>>>(S, out, b->d_frame_off, b->d_block_clip,  \
                                                                          total, rows);                             \
            HPSS_LAUNCHED("median_freq_walk_kernel");                                                               \
        } else {                                                                                                    \
            median_freq_walk_group_kernel<KK><<<grid, kWalkWarps * 32, 0, st>>>(S, out, b->d_frame_off,             \
                                                                                b->d_block_clip, total, rows);      \
            HPSS_LAUNCHED("median_freq_walk_group_kernel");                                                         \
        }                                                                                                           \
        return HPSS_OK;                                                                                             \
    }
    HPSS_MEDIAN_FAST_KS(HPSS_WALK_ANY_K)
#undef HPSS_WALK_ANY_K
    return HPSS_OK;
}

}  // namespace hpss
