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
#include <stdlib.h>

#include <algorithm>

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
__device__ __forceinline__ LineInfo lane_line(const int64_t* __restrict__ frame_off,
                                              const int32_t* __restrict__ block_clip, int rows, int64_t n_lines,
                                              int64_t line, int uniform_T = 0) {
    LineInfo li;
    li.base = 0; li.n = 0; li.estride = 1;
    if (line >= n_lines) return li;
    if (TIME_AXIS && uniform_T > 0) {
        // every clip has uniform_T frames: line (c, f) starts at (c * rows + f) * T = line * T -- no table lookups
        li.base = line * uniform_T;
        li.n = uniform_T;
        return li;
    }
    if (TIME_AXIS) {
        const int c = (int)(line / rows);
        const int f = (int)(line - (int64_t)c * rows);
        const int64_t fo = __ldg(frame_off + c);
        const int T = (int)(__ldg(frame_off + c + 1) - fo);
        li.base = (int64_t)rows * fo + (int64_t)f * T;
        li.n = T;
        li.estride = 1;
    } else {
        const int c = find_clip_hint(frame_off, block_clip, line);
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
        // runs of rows with the same length (see tile_fill_async): one or two per block in practice
        const int nA = __shfl_sync(0xffffffffu, li.n, 0);
        const unsigned sameA = __ballot_sync(0xffffffffu, li.n == nA);
        int rA = 32, rB = 32, nB = nA;
        unsigned rest = 0;
        if (sameA != 0xffffffffu) {
            rA = __ffs(~sameA) - 1;
            nB = __shfl_sync(0xffffffffu, li.n, rA);
            const unsigned sameB = __ballot_sync(0xffffffffu, lane < rA || li.n == nB);
            rB = (sameB == 0xffffffffu) ? 32 : __ffs(~sameB) - 1;
            rest = __ballot_sync(0xffffffffu, lane >= rB && li.n != 0);
        }
        if (rA == 32) {
            // one run: line r starts r * nA elements after line 0
            if (nA == 0 || p0 >= nA) return;
            float* row0 = out + __shfl_sync(0xffffffffu, li.base, 0) + p0;
            const int lim = min(TT, nA - p0);
            const uint32_t pitch = 4u * (uint32_t)nA;      // bytes between rows: one IMAD.WIDE.U32 per address (FMA pipe)
            for (int pos = lane; pos < lim; pos += 32) {
                const float* s = sm + pos;
                char* d0 = reinterpret_cast<char*>(row0 + pos);
#pragma unroll 8
                for (int r = 0; r < 32; ++r, s += lstride)
                    *reinterpret_cast<float*>(d0 + (uint64_t)(uint32_t)r * pitch) = *s;
            }
            return;
        }
        if (rest == 0) {
            // two runs (the block straddles a clip border)
            auto store_run = [&](int rb, int re, int n0) {
                if (rb >= re || n0 == 0 || p0 >= n0) return;
                float* row0 = out + __shfl_sync(0xffffffffu, li.base, rb) + p0;
                const int lim = min(TT, n0 - p0);
                const uint32_t pitch = 4u * (uint32_t)n0;
                for (int pos = lane; pos < lim; pos += 32) {
                    const float* s = sm + rb * lstride + pos;
                    char* d0 = reinterpret_cast<char*>(row0 + pos);
                    for (int r = 0; r < re - rb; ++r, s += lstride)
                        *reinterpret_cast<float*>(d0 + (uint64_t)(uint32_t)r * pitch) = *s;
                }
            };
            store_run(0, rA, nA);
            store_run(rA, rB, nB);
            return;
        }
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
// Persistent CTA = kComputeWarps compute warps + 1 loader warp around a ring of kRing... tile
// buffers.  The loader warp fills tile n (cp.async, reflected at the clip borders) while the
// compute warps run the selection networks on earlier tiles; full/empty mbarriers per buffer.
// Item n of a CTA uses buffer n % NB and is consumed by compute warp n % kComputeWarps.
#ifndef HPSS_COMPUTE_WARPS
#define HPSS_COMPUTE_WARPS 8
#endif
constexpr int kComputeWarps = HPSS_COMPUTE_WARPS;
#ifndef HPSS_LOADER_WARPS
#define HPSS_LOADER_WARPS 4
#endif
constexpr int kLoaderWarps = HPSS_LOADER_WARPS;
constexpr int kRingThreads = (kComputeWarps + kLoaderWarps) * 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async4(uint32_t dst, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
// arrive on `bar` once all cp.async issued so far by this thread have landed (count pre-armed)
__device__ __forceinline__ void cp_async_arrive(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared.b64 [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE;\n"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// loader: fill one tile asynchronously
// (loader warp `lw` of kLoaderWarps takes every kLoaderWarps-th row / position)
struct FillCache {   // time axis, uniform batch: reflected source offsets of this lane's tile positions
    static constexpr int kChunks = 8;
    int p0 = -1, n0 = -1;
    int off[kChunks];
};

// rows r0, r0 + kLoaderWarps, ... < re of one run of equal-length rows: NF whole chunks of 32 positions + one partial
// chunk per row.  Byte address = row + 4 * offset as ONE IMAD.WIDE on the FMA pipe (the multiplier is a run-time 4,
// so the compiler cannot turn it into the two-instruction LEA pair on the ALU pipe, which this kernel saturates).
template <int NF>
__device__ __forceinline__ void fill_rows(const float* rowp, uint32_t drow, int64_t sstep, uint32_t dstep, int r0, int re,
                                          const int (&off)[FillCache::kChunks], uint32_t four, bool tail) {
#pragma unroll 2
    for (int r = r0; r < re; r += kLoaderWarps, drow += dstep, rowp += sstep) {
#pragma unroll
        for (int q = 0; q < NF; ++q)
            cp_async4(drow + 128u * q, reinterpret_cast<const float*>(reinterpret_cast<const char*>(rowp) +
                                                                     (uint64_t)(uint32_t)off[q] * four));
        if (tail)
            cp_async4(drow + 128u * NF, reinterpret_cast<const float*>(reinterpret_cast<const char*>(rowp) +
                                                                      (uint64_t)(uint32_t)off[NF < FillCache::kChunks ? NF : 0] * four));
    }
}

template <bool TIME_AXIS>
__device__ __forceinline__ void tile_fill_async(uint32_t sm_base, const float* __restrict__ S, const LineInfo& li,
                                                int lane, int lw, int p0, int halo, int span, int lstride, FillCache (&fc)[2]) {
    if (TIME_AXIS) {
        // lanes sweep positions of one row at a time (coalesced); reflected source indices are
        // computed once per lane when all 32 rows have the same length (uniform batch)
        // Runs of rows with the same length: the 32 lines of a block are consecutive rows, so rows of equal length
        // are exactly n elements apart (also across clips of equal length).  A block holds one run (inside a
        // clip, or a uniform batch) or two (it straddles a clip border); anything else takes the slow path below.
        const int nA = __shfl_sync(0xffffffffu, li.n, 0);
        const unsigned sameA = __ballot_sync(0xffffffffu, li.n == nA);
        int rA = 32, rB = 32, nB = nA;
        unsigned rest = 0;
        if (sameA != 0xffffffffu) {                             // (warp-uniform) not the common single-run block
            rA = __ffs(~sameA) - 1;                                               // rows [0, rA) have length nA
            nB = __shfl_sync(0xffffffffu, li.n, rA);
            const unsigned sameB = __ballot_sync(0xffffffffu, lane < rA || li.n == nB);
            rB = (sameB == 0xffffffffu) ? 32 : __ffs(~sameB) - 1;                 // rows [rA, rB) have length nB
            rest = __ballot_sync(0xffffffffu, lane >= rB && li.n != 0);
        }
        if (rest == 0) {                                        // at most two runs (+ absent lines after the batch end)
            constexpr int NCH = FillCache::kChunks;             // span = TT + K - 1 <= 16 * 12 + 62 = 254 < 8 * 32
            const uint32_t four = 4u * gridDim.y;               // = 4
            auto fill_run = [&](int rb, int re, int n0, FillCache& c) {
                if (rb >= re || n0 == 0 || p0 >= n0) return;            // warp-uniform
                // The reflected source offsets of this lane's positions depend on (p0, n0) only: they are kept
                // in registers from tile to tile, so a row costs one address add per 128-byte copy.
                if (c.p0 != p0 || c.n0 != n0) {
                    c.p0 = p0; c.n0 = n0;
#pragma unroll
                    for (int q = 0; q < NCH; ++q) c.off[q] = reflect_idx(p0 - halo + lane + 32 * q, n0);
                }
                const int r0 = rb + ((lw - rb) % kLoaderWarps + kLoaderWarps) % kLoaderWarps;   // first row of this warp
                const float* rowp = S + __shfl_sync(0xffffffffu, li.base, rb) + (int64_t)(r0 - rb) * n0;
                uint32_t drow = sm_base + 4u * (uint32_t)(r0 * lstride + lane);
                const uint32_t dstep = 4u * (uint32_t)(kLoaderWarps * lstride);
                const int64_t sstep = (int64_t)kLoaderWarps * n0;
                // whole 32-position chunks unconditionally, the partial chunk under one loop-invariant predicate (a compare
                // per element is an ALU-pipe instruction: see relayout_rows)
                const bool tail = lane < (span & 31);
                switch (span >> 5) {
                    case 0: fill_rows<0>(rowp, drow, sstep, dstep, r0, re, c.off, four, tail); break;
                    case 1: fill_rows<1>(rowp, drow, sstep, dstep, r0, re, c.off, four, tail); break;
                    case 2: fill_rows<2>(rowp, drow, sstep, dstep, r0, re, c.off, four, tail); break;
                    case 3: fill_rows<3>(rowp, drow, sstep, dstep, r0, re, c.off, four, tail); break;
                    case 4: fill_rows<4>(rowp, drow, sstep, dstep, r0, re, c.off, four, tail); break;
                    case 5: fill_rows<5>(rowp, drow, sstep, dstep, r0, re, c.off, four, tail); break;
                    case 6: fill_rows<6>(rowp, drow, sstep, dstep, r0, re, c.off, four, tail); break;
                    default: fill_rows<7>(rowp, drow, sstep, dstep, r0, re, c.off, four, tail); break;
                }
            };
            fill_run(0, rA, nA, fc[0]);
            if (rA < 32) fill_run(rA, rB, nB, fc[1]);
        } else {
            for (int r = lw; r < 32; r += kLoaderWarps) {
                const int64_t b = __shfl_sync(0xffffffffu, li.base, r);
                const int n = __shfl_sync(0xffffffffu, li.n, r);
                if (n == 0 || p0 >= n) continue;
                for (int pos = lane; pos < span; pos += 32)
                    cp_async4(sm_base + 4u * (uint32_t)(r * lstride + pos), S + b + reflect_idx(p0 - halo + pos, n));
            }
        }
    } else {
        if (li.n > 0) {
            const float* col = S + li.base;
            const int64_t es = li.estride;
            // reflected head / tail, plain strided body (no index arithmetic per element)
            const int first = p0 - halo;                       // source index of pos 0
            const int body_lo = max(0, -first);                // first pos with an interior source
            const int body_hi = min(span, li.n - first);       // one past the last interior pos
            for (int pos = lw; pos < min(body_lo, span); pos += kLoaderWarps)
                cp_async4(sm_base + 4u * (uint32_t)(pos * 32 + lane), col + reflect_idx(first + pos, li.n) * es);
            {
                int pos = body_lo + ((lw - body_lo) % kLoaderWarps + kLoaderWarps) % kLoaderWarps;
                const float* src = col + (int64_t)(first + pos) * es;
                uint32_t d = sm_base + 4u * (uint32_t)(pos * 32 + lane);
                const int64_t sstep = es * kLoaderWarps;
#pragma unroll 8
                for (; pos < body_hi; pos += kLoaderWarps, src += sstep, d += 128u * kLoaderWarps) cp_async4(d, src);
            }
            {
                int pos = max(body_hi, 0);
                pos += ((lw - pos) % kLoaderWarps + kLoaderWarps) % kLoaderWarps;
                for (; pos < span; pos += kLoaderWarps)
                    cp_async4(sm_base + 4u * (uint32_t)(pos * 32 + lane), col + reflect_idx(first + pos, li.n) * es);
            }
        }
    }
}

// selection networks of one tile: lane = line, the TT + K - 1 inputs of the line sit at sm[sidx(lane, pos)];
// outputs overwrite the line in place (time axis) or go straight to global memory (frequency axis)
template <int K, bool TIME_AXIS>
__device__ __forceinline__ void median_compute_tile(float* __restrict__ sm, float* __restrict__ out, const LineInfo& li,
                                                    int lane, int p0, int TT, int lstride) {
    constexpr int G = MedianGroup<K>::G;
    constexpr bool STEP = TIME_AXIS && MedianStep<K>::available;
    constexpr int GS = MedianStep<K>::G;
    constexpr bool MIXED = STEP && GS == G;
    const int NG = TT / G;
    const bool live = li.n > 0 && p0 < li.n;
    if (live) {
        const int ng = min(NG, (li.n - p0 + G - 1) / G);
        int g = 0;
        if constexpr (STEP) {
            // stateful walk: two groups per step, the sorted blocks C2, C3 of one step are C0, C1 of the next
            constexpr int NR = MedianStep<K>::NRAW;
            const int live = min(TT, li.n - p0);                       // outputs this line needs
            const int nst = MIXED ? ng / 2 : (live + 2 * GS - 1) / (2 * GS);
            if (nst > 0) {
                float ca[GS], cb[GS];
                {
                    float r0[GS], r1[GS];
#pragma unroll
                    for (int i = 0; i < GS; ++i) {
                        r0[i] = sm[sidx<TIME_AXIS>(lane, GS - 1 + i, lstride)];
                        r1[i] = sm[sidx<TIME_AXIS>(lane, 2 * GS - 1 + i, lstride)];
                    }
                    MedianStep<K>::sort(r0, ca);
                    MedianStep<K>::sort(r1, cb);
                }
                for (int st = 0; st < nst; ++st) {
                    const int b0 = 2 * GS * st;
                    float xr[NR], o[2 * GS], na[GS], nb[GS];
#pragma unroll
                    for (int i = 0; i < NR; ++i)
                        xr[i] = sm[sidx<TIME_AXIS>(lane, b0 + MedianStep<K>::raw_pos(i), lstride)];
                    MedianStep<K>::run(ca, cb, xr, o, na, nb);
#pragma unroll
                    for (int i = 0; i < GS; ++i) { ca[i] = na[i]; cb[i] = nb[i]; }
#pragma unroll
                    for (int j = 0; j < 2 * GS; ++j) sm[sidx<TIME_AXIS>(lane, b0 + j, lstride)] = o[j];
                }
            }
            g = MIXED ? 2 * nst : ng;                                 // !MIXED: the steps covered the tile
        }
        for (; g < ng; ++g) {
            float x[K + G - 1], o[G];
#pragma unroll
            for (int i = 0; i < K + G - 1; ++i) x[i] = sm[sidx<TIME_AXIS>(lane, g * G + i, lstride)];
            MedianGroup<K>::run(x, o);
            if (TIME_AXIS) {
                // in place: a lane's outputs are consecutive in time, the coalesced store needs the
                // transposed view of the tile (tile_store below)
#pragma unroll
                for (int j = 0; j < G; ++j) sm[sidx<TIME_AXIS>(lane, g * G + j, lstride)] = o[j];
            } else {
                // lane = frame: registers -> global memory is already one 128-byte row segment per output
                float* dst = out + li.base + (int64_t)(p0 + g * G) * li.estride;
                const int nj = min(G, li.n - p0 - g * G);
#pragma unroll
                for (int j = 0; j < G; ++j)
                    if (j < nj) dst[(int64_t)j * li.estride] = o[j];
            }
        }
    }
}

// Tile number -> (line block, first position) in a ragged batch.  tile_first[lb] = tiles of the line blocks before lb
// (n_lb + 1 entries).  A warp's tile numbers only grow, so after one binary search it walks the table forwards: one
// global load per line block (~250 tiles of a MUSAN clip), none per tile.  (A list entry per tile cost a dependent
// global load at the top of every loader iteration -- 24 % of the loader warps' time in ncu even when it was issued an
// iteration ahead, because the compiler renames the prefetched register with a move that waits for the load.)
struct TileWalk {
    int64_t lb = -1, first = 0, last = 0;                  // tiles [first, last) belong to line block lb
    __device__ __forceinline__ void locate(const int64_t* __restrict__ tile_first, int64_t n_lb, int64_t item) {
        if (item < last) return;
        if (lb < 0) {                                      // largest lb with tile_first[lb] <= item
            int64_t lo = 0, hi = n_lb - 1;
            while (lo < hi) {
                const int64_t mid = (lo + hi + 1) >> 1;
                if (__ldg(tile_first + mid) <= item) lo = mid; else hi = mid - 1;
            }
            lb = lo;
            first = __ldg(tile_first + lo);
            last = __ldg(tile_first + lo + 1);
        }
        while (item >= last) {                             // (empty line blocks: clips without frames)
            ++lb;
            first = last;
            last = __ldg(tile_first + lb + 1);
        }
    }
};

template <int K, bool TIME_AXIS>
__global__ void __launch_bounds__(kRingThreads, 1)
median_fast_kernel(const float* __restrict__ S, float* __restrict__ out, const int64_t* __restrict__ frame_off,
                   const int32_t* __restrict__ block_clip, int rows, int64_t n_lines, int TT, int n_ptiles,
                   int64_t n_items, int NB, int uniform_T, const int64_t* __restrict__ tile_first) {
    constexpr int G = MedianGroup<K>::G;
    constexpr int HALO = K / 2;
    // stateful double steps where K has them (time axis; the frequency axis has its own walk kernel).  MIXED:
    // same G as the stateless group, which then finishes tiles with an odd number of groups; otherwise tiles
    // are whole steps (launch_fast rounds TT to kGranule<K, TIME_AXIS>)
    constexpr bool STEP = TIME_AXIS && MedianStep<K>::available;
    constexpr int GS = MedianStep<K>::G;
    constexpr bool MIXED = STEP && GS == G;
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int span = TT + K - 1;
    const int lstride = span | 1;
    const int tile_floats = TIME_AXIS ? 32 * lstride : span * 32;
    const int NG = TT / G;
    // barriers live behind the tile ring
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)NB * tile_floats);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + NB);
    if (threadIdx.x == 0) {
        for (int b = 0; b < NB; ++b) {
            mbar_init(full0 + 8u * b, 32 * kLoaderWarps);   // every loader lane arrives once its cp.async land
            mbar_init(empty0 + 8u * b, 1);     // one consumer lane releases the buffer
        }
    }
    __syncthreads();

    // A CTA owns a CONTIGUOUS range of the tile sequence (tiles of one line block are consecutive in it): the line
    // block, hence the per-lane line table (two dependent global loads in a ragged batch), changes once per ~range
    // instead of at nearly every tile as with CTA-strided items, and neighbouring tiles' halos come from L1/L2.
    // (MUSAN-shaped corpus batch, k = 21: 0.68 -> see DESIGN.md ns per frame.)
    const int64_t per_cta = (n_items + gridDim.x - 1) / gridDim.x;
    const int64_t item_first = per_cta * blockIdx.x;
    const int64_t n_lb = (n_lines + 31) >> 5;
    const int64_t my_items = max((int64_t)0, min(per_cta, n_items - item_first));

    if (warp >= kComputeWarps) {
        // ===== loader warps =====
        const int lw = warp - kComputeWarps;
        FillCache fc[2];
        LineInfo li;
        li.base = 0; li.n = 0; li.estride = 1;
        int64_t li_lb = -1;
        // ring slot and parity are running counters: a division / modulo per tile and warp is ~30 ALU-pipe instructions
        int b = 0;
        uint32_t bphase = 1;                               // parity of the buffer's previous release
        bool bfirst = true;
        int64_t item = item_first;
        TileWalk tw;
        for (int64_t n = 0; n < my_items; ++n, ++item) {
            int64_t lb;
            int p0;
            if (tile_first != nullptr) {                   // ragged batch: only the tiles that exist
                tw.locate(tile_first, n_lb, item);
                lb = tw.lb; p0 = (int)(item - tw.first) * TT;
            } else if (n_ptiles == 1) {
                lb = item; p0 = 0;
            } else {
                lb = item / n_ptiles;
                p0 = (int)(item - lb * n_ptiles) * TT;
            }
            if (lb != li_lb) {                             // consecutive tiles of a long clip share the line block
                li = lane_line<TIME_AXIS>(frame_off, block_clip, rows, n_lines, lb * 32 + lane, uniform_T);
                li_lb = lb;
            }
            if (!bfirst) mbar_wait(empty0 + 8u * b, bphase);
            tile_fill_async<TIME_AXIS>(smem_u32(smem + (size_t)b * tile_floats), S, li, lane, lw, p0, HALO, span, lstride, fc);
            cp_async_arrive(full0 + 8u * b);
            if (++b == NB) { b = 0; bphase ^= 1u; bfirst = false; }
        }
    } else {
        // ===== compute warps =====
        LineInfo li;
        li.base = 0; li.n = 0; li.estride = 1;
        int64_t li_lb = -1;
        int b = warp % NB;
        uint32_t phase = (uint32_t)(warp / NB) & 1u;
        int64_t item = item_first + warp;
        TileWalk tw;
        for (int64_t n = warp; n < my_items; n += kComputeWarps, item += kComputeWarps) {
            float* sm = smem + (size_t)b * tile_floats;
            int64_t lb;
            int p0;
            if (tile_first != nullptr) {                   // ragged batch: only the tiles that exist
                tw.locate(tile_first, n_lb, item);
                lb = tw.lb; p0 = (int)(item - tw.first) * TT;
            } else if (n_ptiles == 1) {
                lb = item; p0 = 0;
            } else {
                lb = item / n_ptiles;
                p0 = (int)(item - lb * n_ptiles) * TT;
            }
            if (lb != li_lb) {
                li = lane_line<TIME_AXIS>(frame_off, block_clip, rows, n_lines, lb * 32 + lane, uniform_T);
                li_lb = lb;
            }
            mbar_wait(full0 + 8u * b, phase);
            median_compute_tile<K, TIME_AXIS>(sm, out, li, lane, p0, TT, lstride);
            __syncwarp();
            if (TIME_AXIS) {
                tile_store<TIME_AXIS>(sm, out, li, lane, p0, TT, lstride);
                __syncwarp();
            }
            if (lane == 0) mbar_arrive(empty0 + 8u * b);
            b += kComputeWarps;
            while (b >= NB) { b -= NB; phase ^= 1u; }
        }
    }
}

// ---- dense path (time axis, batches of equal clips whose whole line fits one tile: the training-segment shape) -----
// The 32 lines of a tile are 32 consecutive rows of T floats = ONE contiguous, 16-byte aligned chunk of 128 * T bytes,
// so a single elected thread fetches the whole tile with one bulk asynchronous copy (cp.async.bulk, completion counted
// in bytes on an mbarrier) into a small ring of dense landing buffers.  The loader warps then only re-lay the tile
// inside shared memory -- dense rows -> rows with the reflected halo and an odd stride (LDS + STS, the reflected
// source offsets of a lane's positions live in registers) -- instead of gathering it from global memory with
// 4-byte cp.async (one copy per ~75 cycles and warp: the loader ring was the limit of this kernel for k <= 31).
// Compute warps, selection networks and the coalesced store are those of median_fast_kernel.
#ifndef HPSS_DENSE_MINK
#define HPSS_DENSE_MINK 3     // smallest kernel size that takes the dense bulk-copy path (all of them since the loader lost its compares)
#endif
#ifndef HPSS_LAND
#define HPSS_LAND 3
#endif
#ifndef HPSS_RELAYOUT_FULL
#define HPSS_RELAYOUT_FULL 1
#endif
constexpr int kLand = HPSS_LAND;  // landing buffers

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];\n" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;\n" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}

// dense rows -> padded rows of one tile for the rows lw, lw + kLoaderWarps, ...: NF whole chunks of 32 positions
// plus one partial chunk (`tail` = this lane takes part in it); span <= 32 * FillCache::kChunks
template <int NF>
__device__ __forceinline__ void relayout_row(uint32_t src, uint32_t dst, const uint32_t (&off4)[FillCache::kChunks],
                                             bool tail) {
    float v[NF + 1];
#pragma unroll
    for (int q = 0; q < NF; ++q) v[q] = lds_f32(src + off4[q]);
    if (tail) v[NF] = lds_f32(src + off4[NF < FillCache::kChunks ? NF : 0]);
#pragma unroll
    for (int q = 0; q < NF; ++q) sts_f32(dst + 128u * q, v[q]);
    if (tail) sts_f32(dst + 128u * NF, v[NF]);
}
template <int NF>
__device__ __forceinline__ void relayout_rows(uint32_t src, uint32_t dst, uint32_t sstep, uint32_t dstep, int lw,
                                              int rows_here, const uint32_t (&off4)[FillCache::kChunks], bool tail) {
    if (HPSS_RELAYOUT_FULL && rows_here == 32) {              // every tile but the last: no loop counter, no compare (ALU pipe)
#pragma unroll
        for (int i = 0; i < 32 / kLoaderWarps; ++i) relayout_row<NF>(src + i * sstep, dst + i * dstep, off4, tail);
        return;
    }
    for (int r = lw; r < rows_here; r += kLoaderWarps, src += sstep, dst += dstep) relayout_row<NF>(src, dst, off4, tail);
}

template <int K>
__global__ void __launch_bounds__(kRingThreads, 1)
median_dense_kernel(const float* __restrict__ S, float* __restrict__ out, int64_t n_lines, int T, int TT, int64_t n_items,
                    int NB) {
    constexpr int HALO = K / 2;
    extern __shared__ __align__(128) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int span = TT + K - 1;
    const int lstride = span | 1;
    const int tile_floats = 32 * lstride;
    const int land_floats = (32 * T + 31) & ~31;                  // 128-byte multiples keep every buffer 16-byte aligned
    float* land0 = smem + (((size_t)NB * tile_floats + 31) & ~(size_t)31);     // the base is 128-byte aligned
    uint64_t* bars = reinterpret_cast<uint64_t*>(land0 + (size_t)kLand * land_floats);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + NB);
    const uint32_t lfull0 = smem_u32(bars + 2 * NB), lempty0 = smem_u32(bars + 2 * NB + kLand);
    if (threadIdx.x == 0) {
        for (int b = 0; b < NB; ++b) {
            mbar_init(full0 + 8u * b, 32 * kLoaderWarps);         // every loader lane arrives after its stores
            mbar_init(empty0 + 8u * b, 1);                        // one consumer lane releases the buffer
        }
        for (int b = 0; b < kLand; ++b) {
            mbar_init(lfull0 + 8u * b, 1);                        // the issuing thread's expect_tx arrival + the bytes
            mbar_init(lempty0 + 8u * b, kLoaderWarps);            // every loader warp has read the landing buffer
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    // items of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...  Ring slots and phases are running counters (a
    // division / modulo per tile and warp is ~30 ALU-pipe instructions); the launcher keeps the count below 2^31.
    const int my_items = (n_items > blockIdx.x) ? (int)((n_items - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    auto tile_rows = [&](int64_t item) { return (int)min((int64_t)32, n_lines - item * 32); };

    if (warp >= kComputeWarps) {
        // ===== loader warps =====
        const int lw = warp - kComputeWarps;
        // reflected source BYTE offset (within a dense row) of this lane's positions; (p0, n) = (0, T) for every tile
        uint32_t off4[FillCache::kChunks];
#pragma unroll
        for (int q = 0; q < FillCache::kChunks; ++q) off4[q] = 4u * (uint32_t)reflect_idx(lane + 32 * q - HALO, T);
        int islot = 0;                                             // landing slot / its use parity of the next issue
        uint32_t iphase = 1;                                       // (parity of use - 1: first round never waits)
        int64_t iitem = blockIdx.x;
        bool ifirst = true;
        auto issue = [&]() {                                       // elected thread: bulk copy of the next item
            if (!ifirst) mbar_wait(lempty0 + 8u * islot, iphase);
            const uint32_t bytes = ((uint32_t)tile_rows(iitem) * (uint32_t)T * 4u) & ~15u;
            mbar_expect_tx(lfull0 + 8u * islot, bytes);
            if (bytes) bulk_load(smem_u32(land0 + (size_t)islot * land_floats), S + iitem * 32 * (int64_t)T, bytes, lfull0 + 8u * islot);
            iitem += gridDim.x;
            if (++islot == kLand) { islot = 0; iphase ^= 1u; ifirst = false; }
        };
        const bool elected = lw == 0 && lane == 0;
        if (elected)
            for (int n = 0; n < min(kLand - 1, my_items); ++n) issue();
        int slot = 0, b = 0;
        uint32_t lphase = 0, bphase = 1;                           // parity to wait for: landing full / tile empty
        bool bfirst = true;
        int64_t item = blockIdx.x;
        const uint32_t land_b = smem_u32(land0), tile_b = smem_u32(smem);
        for (int n = 0; n < my_items; ++n, item += gridDim.x) {
            if (elected && n + kLand - 1 < my_items) issue();
            const int rows_here = tile_rows(item);
            const uint32_t land = land_b + 4u * (uint32_t)(slot * land_floats);
            mbar_wait(lfull0 + 8u * slot, lphase);
            if (rows_here != 32) {   // the (at most three) floats behind the last 16-byte unit of a partial last tile
                const int got = (int)((((uint32_t)rows_here * (uint32_t)T * 4u) & ~15u) / 4u), want = rows_here * T;
                if (lw == 0 && got + lane < want)
                    sts_f32(land + 4u * (uint32_t)(got + lane), __ldg(S + item * 32 * (int64_t)T + got + lane));
                if (got != want) asm volatile("bar.sync 1, %0;\n" ::"n"(32 * kLoaderWarps) : "memory");
            }
            if (!bfirst) mbar_wait(empty0 + 8u * b, bphase);
            // 32-bit shared-memory addresses: one add per access (generic pointers cost four ALU instructions each)
            const uint32_t src = land + 4u * (uint32_t)(lw * T);
            const uint32_t dst = tile_b + 4u * (uint32_t)(b * tile_floats + lw * lstride + lane);
            const uint32_t sstep = 4u * (uint32_t)(kLoaderWarps * T), dstep = 4u * (uint32_t)(kLoaderWarps * lstride);
            // (four rows per iteration, all loads before the first store, was measured slower: 0.32 vs 0.24 ms at k = 31)
            // The number of whole 32-position chunks of a row is a compile-time constant of the copy loop (dispatched
            // once per tile) and the predicate of the partial chunk is loop invariant: no compare per element (ISETP
            // runs on the ALU pipe, which the selection networks saturate; the predicated version spent 9 % of the
            // kernel's ALU-pipe time on them).
            const bool tail = lane < (span & 31);
            switch (span >> 5) {
                case 0: relayout_rows<0>(src, dst, sstep, dstep, lw, rows_here, off4, tail); break;
                case 1: relayout_rows<1>(src, dst, sstep, dstep, lw, rows_here, off4, tail); break;
                case 2: relayout_rows<2>(src, dst, sstep, dstep, lw, rows_here, off4, tail); break;
                case 3: relayout_rows<3>(src, dst, sstep, dstep, lw, rows_here, off4, tail); break;
                case 4: relayout_rows<4>(src, dst, sstep, dstep, lw, rows_here, off4, tail); break;
                case 5: relayout_rows<5>(src, dst, sstep, dstep, lw, rows_here, off4, tail); break;
                case 6: relayout_rows<6>(src, dst, sstep, dstep, lw, rows_here, off4, tail); break;
                default: relayout_rows<7>(src, dst, sstep, dstep, lw, rows_here, off4, tail); break;
            }
            mbar_arrive(full0 + 8u * b);
            __syncwarp();
            if (lane == 0) mbar_arrive(lempty0 + 8u * slot);
            if (++slot == kLand) { slot = 0; lphase ^= 1u; }
            if (++b == NB) { b = 0; bphase ^= 1u; bfirst = false; }
        }
    } else {
        // ===== compute warps =====
        int b = warp;                                              // NB > kComputeWarps: at most one wrap per item
        uint32_t phase = 0;
        int64_t item = blockIdx.x + (int64_t)warp * gridDim.x;
        for (int n = warp; n < my_items; n += kComputeWarps, item += (int64_t)kComputeWarps * gridDim.x) {
            float* sm = smem + (size_t)b * tile_floats;
            LineInfo li;
            li.estride = 1;
            li.n = (item * 32 + lane < n_lines) ? T : 0;
            li.base = (item * 32 + lane) * (int64_t)T;
            mbar_wait(full0 + 8u * b, phase);
            median_compute_tile<K, true>(sm, out, li, lane, 0, TT, lstride);
            __syncwarp();
            // (a store with every address formed on the FMA pipe -- fma_pipe_store, rows unrolled 4 / 8 / 32 times --
            // was measured slower than this one at every kernel size: 0.224 - 0.241 vs 0.220 ms at k = 31)
            tile_store<true>(sm, out, li, lane, 0, TT, lstride);
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8u * b);
            b += kComputeWarps;
            if (b >= NB) { b -= NB; phase ^= 1u; }
        }
    }
}

// ---- generic path: any k (even k, k > 63): rank counting, O(k^2) per output ---------
template <bool TIME_AXIS>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
median_generic_kernel(const float* __restrict__ S, float* __restrict__ out, const int64_t* __restrict__ frame_off,
                      const int32_t* __restrict__ block_clip, int rows, int64_t n_lines, int k, int TT, int n_ptiles,
                      int64_t n_items) {
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
        const LineInfo li = lane_line<TIME_AXIS>(frame_off, block_clip, rows, n_lines, lb * 32 + lane);
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

// tiles of a ragged batch: only those that exist, as a prefix count per line block (tile -> (line block, first position)
// by TileWalk in the kernel); cached in the batch
static int time_tile_list(hpss_ctx* ctx, const hpss_batch* cb, int rows, int TT, const int64_t** d_first, int64_t* n_tiles) {
    hpss_batch* b = const_cast<hpss_batch*>(cb);
    std::lock_guard<std::mutex> lk(ctx->mu);
    const auto key = std::make_pair(rows, TT);
    auto it = b->time_tiles.find(key);
    if (it == b->time_tiles.end()) {
        const int64_t n_lines = (int64_t)b->n_clips * rows;
        const int64_t n_lb = (n_lines + 31) / 32;
        std::vector<int64_t> first((size_t)n_lb + 1, 0);        // first[lb] = tiles of the line blocks before lb
        for (int64_t lb = 0; lb < n_lb; ++lb) {
            const int c0 = (int)((lb * 32) / rows);
            const int c1 = (int)(std::min(n_lines - 1, lb * 32 + 31) / rows);
            int64_t tmax = 0;
            for (int c = c0; c <= c1; ++c) tmax = std::max(tmax, b->frame_off[c + 1] - b->frame_off[c]);
            first[lb + 1] = first[lb] + (tmax + TT - 1) / TT;
        }
        int64_t* d = nullptr;
        HPSS_CUDA(cudaMalloc(&d, sizeof(int64_t) * first.size()));
        HPSS_CUDA(cudaMemcpy(d, first.data(), sizeof(int64_t) * first.size(), cudaMemcpyHostToDevice));
        it = b->time_tiles.emplace(key, std::make_pair(d, first.back())).first;
    }
    *d_first = it->second.first;
    *n_tiles = it->second.second;
    return HPSS_OK;
}

template <int K, bool TIME_AXIS>
int launch_fast(hpss_ctx* ctx, const hpss_batch* b, const float* S, float* out, const int64_t* d_frame_off,
                const int32_t* d_block_clip, int rows, int64_t n_lines, int64_t max_len, int uniform_T, cudaStream_t st) {
    // tile length granule: a stateless group, or a whole stateful double step when its G differs from the group's
    constexpr int G = (TIME_AXIS && MedianStep<K>::available && MedianStep<K>::G != MedianGroup<K>::G)
                          ? 2 * MedianStep<K>::G : MedianGroup<K>::G;
    // largest tile (multiple of G outputs, at most 16 groups) that still leaves room for a ring of
    // kComputeWarps + 2 buffers in the opt-in shared memory
    const size_t per_buf = ((size_t)ctx->max_smem_optin - 512) / (kComputeWarps + 2) - 16;
    const int max_span = (int)(per_buf / (32 * sizeof(float))) - 1;
    int tt_max = (max_span - (K - 1)) / G * G;
    const int tt_cap = std::max(16 * G, 128 / G * G);      // small kernels have small groups: still ~128 outputs per tile
    if (tt_max > tt_cap) tt_max = tt_cap;
    if (tt_max < G) {
        set_error("median k=%d does not fit the shared-memory tile ring", K);
        return HPSS_ERR_UNSUPPORTED;
    }
    const int64_t nt0 = (max_len + tt_max - 1) / tt_max;
    int TT = (int)((max_len + nt0 - 1) / nt0);
    TT = (TT + G - 1) / G * G;
    const int n_ptiles = (int)((max_len + TT - 1) / TT);
    const int64_t n_lb = (n_lines + 31) / 32;
    int64_t n_items = n_lb * n_ptiles;
    const int64_t* tile_list = nullptr;
    if (TIME_AXIS && uniform_T == 0 && n_ptiles > 1) {
        // ragged batch: only the tiles that exist (clips shorter than the longest have fewer)
        const int rc = time_tile_list(ctx, b, rows, TT, &tile_list, &n_items);
        if (rc) return rc;
        if (n_items == 0) return HPSS_OK;
    }
    const int span = TT + K - 1;
    const size_t tile_bytes = (TIME_AXIS ? (size_t)32 * (span | 1) : (size_t)span * 32) * sizeof(float);
    if constexpr (TIME_AXIS) {
        // equal clips whose whole line fits one tile, 16-byte aligned input: one bulk copy per tile (median_dense_kernel)
        // (measured on the 4096 x 1 s batch: k = 31 0.248 -> 0.242 ms, k = 41 0.351 -> 0.339 ms, but k = 21 0.224 -> 0.231 ms
        // and k = 11 0.201 -> 0.212 ms: below ~25 taps the tile period is too short for the two-stage pipeline)
        if (K >= HPSS_DENSE_MINK && uniform_T > 0 && n_ptiles == 1 && span < 32 * FillCache::kChunks && !knobs().no_dense_median &&
            (reinterpret_cast<uintptr_t>(S) & 15) == 0) {
            const size_t land_bytes = (((size_t)32 * uniform_T + 31) & ~(size_t)31) * sizeof(float);
            const size_t fixed = kLand * land_bytes + 128 + 16 * (size_t)kLand + 256;
            int NBd = (int)(((size_t)ctx->max_smem_optin - fixed) / (tile_bytes + 16));
            if (NBd > 2 * kComputeWarps) NBd = 2 * kComputeWarps;
            if (NBd >= kComputeWarps + 1) {
                const size_t smem_d = (size_t)NBd * tile_bytes + fixed + (size_t)NBd * 16;
                auto kd = median_dense_kernel<K>;
                HPSS_CUDA(cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_d));
                int64_t gd = (n_lb + kComputeWarps - 1) / kComputeWarps;
                if (gd > ctx->sm_count) gd = ctx->sm_count;
                if (n_lb / gd >= 0x7fffffff) { set_error("median: batch too large"); return HPSS_ERR_UNSUPPORTED; }
                kd<<<(unsigned)gd, kRingThreads, smem_d, st>>>(S, out, n_lines, uniform_T, TT, n_lb, NBd);
                HPSS_LAUNCHED("median_dense_kernel");
                return HPSS_OK;
            }
        }
    }
    int NB = (int)(((size_t)ctx->max_smem_optin - 512) / (tile_bytes + 16));
    if (NB > 2 * kComputeWarps) NB = 2 * kComputeWarps;
    const size_t smem = (size_t)NB * tile_bytes + (size_t)NB * 16;
    auto kern = median_fast_kernel<K, TIME_AXIS>;
    HPSS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t grid = (n_items + kComputeWarps - 1) / kComputeWarps;
    if (grid > ctx->sm_count) grid = ctx->sm_count;
    kern<<<(unsigned)grid, kRingThreads, smem, st>>>(S, out, d_frame_off, d_block_clip, rows, n_lines, TT, n_ptiles,
                                                     n_items, NB, uniform_T, tile_list);
    HPSS_LAUNCHED("median_fast_kernel");
    return HPSS_OK;
}

template <bool TIME_AXIS>
int launch_generic(hpss_ctx* ctx, const float* S, float* out, const int64_t* d_frame_off, const int32_t* d_block_clip, int rows,
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
    kern<<<(unsigned)grid, kWarpsPerCta * 32, smem, st>>>(S, out, d_frame_off, d_block_clip, rows, n_lines, k, TT,
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
    const int uniform_T = (b->uniform_frames > 0 && b->uniform_frames < 0x7fffffff) ? (int)b->uniform_frames : 0;
    if (k == 1) {
        HPSS_CUDA(cudaMemcpyAsync(out, S, sizeof(float) * (size_t)rows * total_frames, cudaMemcpyDeviceToDevice, st));
        return HPSS_OK;
    }
    if (!time_axis) {
        // frequency axis: register walk, no shared-memory staging (median_walk.cu), every generated kernel size
        bool handled = false;
        const int rc = launch_median_freq_walk(ctx, b, S, rows, k, out, st, &handled);
        if (rc || handled) return rc;
    } else {
#define HPSS_DISPATCH_K(KK)                                                                                   \
        if (k == KK)                                                                                          \
            return launch_fast<KK, true>(ctx, b, S, out, b->d_frame_off, b->d_block_clip, rows, n_lines, max_len, \
                                         uniform_T, st);
        HPSS_MEDIAN_FAST_KS(HPSS_DISPATCH_K)
#undef HPSS_DISPATCH_K
    }
    return time_axis ? launch_generic<true>(ctx, S, out, b->d_frame_off, b->d_block_clip, rows, n_lines, max_len, k, st)
                     : launch_generic<false>(ctx, S, out, b->d_frame_off, b->d_block_clip, rows, n_lines, max_len, k, st);
}

}  // namespace hpss
