// Internal declarations shared by the kernels and the C-ABI layer of libhpss_b200.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "../../include/hpss_b200.h"

namespace hpss {

constexpr int kSMs = 148;               // B200: 2 dies x 74 SMs
constexpr int kMaxRadixPasses = 16;

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
extern std::atomic<uint64_t> g_launches;

#define HPSS_CUDA(x)                                              \
    do {                                                          \
        cudaError_t e_ = (x);                                     \
        if (e_ != cudaSuccess) return ::hpss::cuda_fail(e_, #x);  \
    } while (0)

#define HPSS_LAUNCHED(name)                                           \
    do {                                                              \
        ::hpss::g_launches.fetch_add(1, std::memory_order_relaxed);   \
        cudaError_t e_ = cudaGetLastError();                          \
        if (e_ != cudaSuccess) return ::hpss::cuda_fail(e_, name);    \
    } while (0)

// ---- plans ------------------------------------------------------------------------
struct FftPlan {
    int n_fft = 0, win = 0, n2 = 0;
    int n_pass = 0;
    int radix[kMaxRadixPasses] = {0};
    float* d_window = nullptr;   // n_fft floats (periodic Hann, centre padded)
    float* d_window_half = nullptr;   // the same x 0.5 (stft_fast.cu)
    float2* d_tw_half = nullptr; // n2 entries exp(-2 pi i k / n2)
    float2* d_tw_full = nullptr; // n2+1 entries exp(-2 pi i k / n_fft)
    // stft_fast.cu (n2 = NA x NB): tables laid out in the order the two passes read them, so that one 16-byte
    // load fetches two entries:  win_bq[b*NA + q] = 0.5 * (w[2n], w[2n+1]), n = NB*q + b;  tw_kb[k1*NB + b] = W_n2^(b*k1)
    float2* d_win_bq = nullptr;
    float2* d_tw_kb = nullptr;
    // stft_r400_kernel: win_r400[b*20 + q] = w[20q + b];  tw_r400[k1*20 + b] = W_400^(b*k1), k1 = 0..10
    float* d_win_r400 = nullptr;
    float2* d_tw_r400 = nullptr;
};

struct MelPlan {
    int sr = 0, n_fft = 0, n_mels = 0, rows = 0;
    float* d_w = nullptr;        // dense (n_mels, rows)
    int2* d_band = nullptr;      // per filter [first, last+1) non-zero column
    // per frequency row {lowest unfinished filter, its weight, weight of the next filter, 0}: lets a
    // sweep along f keep only two running sums per stream (valid when <= 2 ordered filters overlap)
    int4* d_sweep = nullptr;
    bool sweepable = false;
    // the same sweep split for the fused walk kernel: filters finishing before row f (4 bits per row, 8 rows
    // per word) and the two weights of row f; valid when no row finishes more than 15 filters
    uint32_t* d_emit4 = nullptr;
    float2* d_sweep_w = nullptr;
    bool walkable = false;
};

}  // namespace hpss

struct hpss_ctx {
    int device = 0;
    std::mutex mu;
    std::map<std::pair<int, int>, hpss::FftPlan*> fft_plans;
    std::map<std::tuple<int, int, int>, hpss::MelPlan*> mel_plans;
    std::map<std::pair<int, int>, float*> dct_plans;       // (M, n_mfcc) -> device D^T (M x 4*ceil(n_mfcc/4))
    // grow-only device workspace (S, harm, perc, clip_max, band scratch)
    void* ws = nullptr;
    size_t ws_bytes = 0;
    int2* band_scratch = nullptr;
    int band_scratch_n = 0;
    // host pipeline
    cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr;
    void* pipe_dev[2] = {nullptr, nullptr};   // double-buffered wave + feature staging
    size_t pipe_bytes = 0;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
    int max_smem_optin = 0;
    int sm_count = hpss::kSMs;
    // signal-preparation scratch (prep.cu): descriptors, chunk partials, per-frame gate arrays
    void* prep_ws = nullptr;
    size_t prep_ws_bytes = 0;
    // Every entry point that touches the context's scratch (ws, prep_ws, band_scratch, pipe slots) holds ws_mu
    // while it enqueues, makes its stream wait for ws_done first and records ws_done when it is through: calls
    // from different streams or threads are serialised on the device instead of overwriting each other's S / harm /
    // perc (see hpss::ScratchLease).
    std::mutex ws_mu;
    cudaEvent_t ws_done = nullptr;
    // device status word: bit 0 non-finite audio, bit 1 negative spectrogram input; read by hpss_ctx_check
    uint32_t* d_flags = nullptr;
    uint32_t* h_flags = nullptr;      // pinned mirror
};

struct hpss_batch {
    hpss_ctx* ctx = nullptr;
    int n_clips = 0;
    int n_fft = 0, hop = 0;
    bool has_samples = false;
    std::vector<int64_t> sample_off;   // n_clips + 1
    std::vector<int64_t> frame_off;    // n_clips + 1
    int64_t max_frames = 0;
    int64_t uniform_frames = 0;        // T when every clip has exactly T frames, else 0
    int64_t uniform_samples = 0;       // L when every clip has exactly L samples, else 0
    int64_t* d_sample_off = nullptr;
    int64_t* d_frame_off = nullptr;
    int32_t* d_block_clip = nullptr;   // clip of the first frame of every 32-frame block of the batch
    // K1 tile list (clip, first frame), built for a given tile height
    int stft_tt = 0;
    int n_stft_tiles = 0;
    int2* d_stft_tiles = nullptr;
    int32_t* d_clip_class = nullptr;   // class per clip for the moment kernels (kept across calls: re-uploaded only
    std::vector<int32_t> h_clip_class; //   when the caller passes different classes)
    std::mutex mu;                     // guards the lazily built members below
    struct hpss_pipeline* host_pipe = nullptr;   // cached pipeline of hpss_featuregram_host
    // patch offsets per (patch_size, patch_shift) for hpss_patch_tensor: host prefix + device copy
    std::map<std::pair<int, int>, std::pair<std::vector<int64_t>, int64_t*>> patch_offs;
    // K2h tile lists of ragged batches, one per (rows, tile length): the (line block, first position) pairs that
    // actually exist (a grid over the longest clip would be mostly empty for MUSAN-shaped length distributions)
    std::map<std::pair<int, int>, std::pair<int64_t*, int64_t>> time_tiles;   // (rows, TT) -> tile prefix per line block, tiles
};

namespace hpss { struct PrepPlan; }

// Host-buffer pipeline (api.cu): clip chunks, their sub-batches and the double-buffered device slots
struct hpss_pipeline {
    hpss_ctx* ctx = nullptr;
    hpss_params prm{};
    int n_clips = 0;
    int pcm_format = 0;
    int prepare = 0, fs = 16000;
    double alpha = 0.025, beta = 0.075;
    int rows_out = 0;
    std::vector<int64_t> in_len, in_off;       // raw input samples per clip / prefix
    std::vector<int64_t> wav_len;              // samples per clip the STFT sees (after the preparation)
    std::vector<int64_t> frame_off;            // prefix of STFT frames
    std::vector<int> cut;                      // chunk i = clips [cut[i], cut[i+1])
    std::vector<hpss_batch*> subs;
    std::vector<hpss::PrepPlan*> prep_plans;   // per chunk, when prepare
    size_t in_bytes = 0, wav_bytes = 0, out_bytes = 0;    // per-slot capacities (largest chunk)
    void* slot[2] = {nullptr, nullptr};
    double* d_acc = nullptr;                   // device moment accumulator
    double* h_acc = nullptr;                   // pinned mirror
    int acc_n = 0, acc_classes = 0;
    int32_t* d_class = nullptr;                // class of every clip of the pipeline
    size_t ws_need = 0;
};

namespace hpss {

int get_fft_plan(hpss_ctx* ctx, int n_fft, int win, FftPlan** out);
bool stft_fast_split(int n_fft, int* na, int* nb);      // the NA x NB split of the specialised kernel, if any
int get_mel_plan(hpss_ctx* ctx, int sr, int n_fft, int n_mels, MelPlan** out);
int ensure_workspace(hpss_ctx* ctx, size_t bytes);

// Serialises the users of the context scratch across streams (see hpss_ctx::ws_done).
struct ScratchLease {
    hpss_ctx* ctx = nullptr;
    cudaStream_t st = nullptr;
    std::unique_lock<std::mutex> lk;
    int begin(hpss_ctx* c, cudaStream_t s) {
        lk = std::unique_lock<std::mutex>(c->ws_mu);
        ctx = c; st = s;
        cudaError_t e = cudaStreamWaitEvent(s, c->ws_done, 0);
        if (e != cudaSuccess) return cuda_fail(e, "cudaStreamWaitEvent(ws_done)");
        return HPSS_OK;
    }
    ~ScratchLease() { if (ctx) cudaEventRecord(ctx->ws_done, st); }
};

// environment knobs (development only), read once per process
struct Knobs {
    int no_sweep, host_chunks, no_uniform_moments, mom_ctas, no_fast_stft, no_uniform_stft, k1_real, sweep1, sweep_u,
        dct_grid_mult, no_dense_median;
};
const Knobs& knobs();

// signal preparation (prep.cu)
int64_t prep_out_length(int64_t n, int fs);
int64_t prep_num_frames(int64_t n, int win, int hop);
int launch_prep(hpss_ctx* ctx, const void* pcm, int pcm_format, const int64_t* clip_len, int n_clips, int fs, int win,
                int hop, double alpha, double beta, float* out, int32_t* frame_marker, uint8_t* sample_marker,
                int32_t* n_sil, cudaStream_t st);
struct PrepPlan {                  // prebuilt device descriptors of one batch of files (prep.cu)
    int n_clips = 0;
    size_t n_chunks = 0, n_tiles = 0, n_fr = 0, work_bytes = 0;
    void* d_desc = nullptr;
};
int prep_plan_build(hpss_ctx* ctx, const int64_t* clip_len, int n_clips, int fs, int win, int hop, PrepPlan** out);
void prep_plan_free(PrepPlan* pp);
int launch_prep_plan(hpss_ctx* ctx, const PrepPlan* pp, const void* pcm, int pcm_format, int fs, int win, int hop,
                     double alpha, double beta, float* out, cudaStream_t st);
// mode 0: non-finite -> status bit 0; mode 1: negative -> status bit 1 (read by hpss_ctx_check)
int launch_check(hpss_ctx* ctx, const float* x, int64_t n, int mode, cudaStream_t st);
int launch_mix(hpss_ctx* ctx, const float* sp, const int64_t* sp_len, const float* mu, const int64_t* mu_len,
               const double* target_db, int n_pairs, float* out, cudaStream_t st);
int ensure_stft_tiles(hpss_batch* b, int tt);

// host-side table builders (double precision)
int build_mel(int sr, int n_fft, int n_mels, float* out);
void build_window(int n_fft, int win, float* out);

// stage launchers (device pointers, enqueue on stream)
int launch_stft(hpss_ctx* ctx, hpss_batch* b, const float* wave, const FftPlan* plan, int hop,
                int power, float* S, float* cplx, cudaStream_t st);
int launch_stft_fast(hpss_ctx* ctx, hpss_batch* b, const float* wave, const FftPlan* plan, int hop, int power,
                     float* S, float* cplx, cudaStream_t st, bool* handled);
int launch_median(hpss_ctx* ctx, const hpss_batch* b, const float* S, int rows, int k, bool time_axis,
                  float* out, cudaStream_t st);
int launch_mask_mel(hpss_ctx* ctx, const hpss_batch* b, const float* S, const float* harm,
                    const float* perc, int rows, const float* mel, const int2* band, const int4* sweep,
                    const uint32_t* emit4, const float2* sweep_w, int n_mels, int pre_square, int log_power,
                    float amin, float* out, uint32_t* clip_max, cudaStream_t st);
int launch_median_freq_walk(hpss_ctx* ctx, const hpss_batch* b, const float* S, int rows, int k, float* out,
                            cudaStream_t st, bool* handled);
int launch_mel_bands(const float* mel, int n_mels, int rows, int2* band, cudaStream_t st);
int launch_topdb(hpss_ctx* ctx, const hpss_batch* b, float* out, int rows_per_stream, int n_streams,
                 const uint32_t* clip_max, float top_db, cudaStream_t st);
int launch_moments(hpss_ctx* ctx, const hpss_batch* b, const float* feat, int D, const int32_t* d_class,
                   int n_classes, double* sum, double* sumsq, double* count, double* nonfinite,
                   cudaStream_t st);
int launch_topdb_moments(hpss_ctx* ctx, const hpss_batch* b, float* feat, int rows_per_stream, int n_streams,
                         const uint32_t* clip_max, float top_db, const int32_t* d_class, int n_classes, double* sum,
                         double* sumsq, double* count, double* nonfinite, cudaStream_t st);
int launch_dct(hpss_ctx* ctx, const hpss_batch* b, const float* feat, int M, int n_streams, int n_mfcc, float* out,
               cudaStream_t st);
void build_dct_basis_t(int M, int n_mfcc, int ncols, float* out);
int launch_scale(hpss_ctx* ctx, const hpss_batch* b, const float* feat, int D, const float* mean,
                 const float* stdev, double eps, double* out, float* out32, cudaStream_t st);
int launch_row_standardize(hpss_ctx* ctx, const hpss_batch* b, float* feat, int D, cudaStream_t st);
int launch_patches(const float* feat, int D, int64_t T, int W, int shift, int64_t n_patches, double* out,
                   cudaStream_t st);
int launch_row_nonfinite(hpss_ctx* ctx, const hpss_batch* b, const float* feat, int D, uint8_t* flags, cudaStream_t st);
int launch_patch_tensor(const hpss_batch* b, const void* feat, int in_f64, const int64_t* d_patch_off, int64_t n_patches,
                        int D, int row0, int n_rows, int W, int shift, int time_major, int out_f64, void* out,
                        cudaStream_t st);
int launch_patch_stats(hpss_ctx* ctx, const double* x, int64_t N, int A, int B, int stat, int along_a, double* out,
                       cudaStream_t st);

// ---- device helpers ---------------------------------------------------------------
#ifdef __CUDACC__
// scipy.ndimage 'reflect' (half-sample symmetric), valid for any overshoot.
__device__ __forceinline__ int reflect_idx(int i, int n) {
    if ((unsigned)i < (unsigned)n) return i;          // interior: no reflection
    if (i < 0 && i >= -n) return -1 - i;              // one reflection at the left border
    if (i >= n && i < 2 * n) return 2 * n - 1 - i;    // one reflection at the right border
    const int p = 2 * n;
    i %= p;
    if (i < 0) i += p;
    return i < n ? i : p - 1 - i;
}

// largest c with off[c] <= g  (off has n+1 entries, off[0] == 0, off[n] == total)
__device__ __forceinline__ int find_clip(const int64_t* __restrict__ off, int n, int64_t g) {
    int lo = 0, hi = n;   // invariant: off[lo] <= g < off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(off + mid) <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// clip of global frame g, starting from the clip of its 32-frame block (block_clip[g / 32]):
// a short forward walk instead of a binary search through global memory
__device__ __forceinline__ int find_clip_hint(const int64_t* __restrict__ off, const int32_t* __restrict__ block_clip,
                                              int64_t g) {
    int c = __ldg(block_clip + (g >> 5));
    while (__ldg(off + c + 1) <= g) ++c;
    return c;
}

// ---- addressing without the ALU pipe ------------------------------------------------------------------------------
// The median kernels saturate the ALU pipe with FMNMX (half rate); every IADD3 / VIADD / LEA / ISETP of the address
// arithmetic around them takes a slot of that pipe.  These helpers keep the address arithmetic on the FMA pipe:
//   fma_pipe_mad    a * b + c in 32 bits as one IMAD (PTX, so that the compiler cannot strength-reduce a row loop into
//                   pointer increments = IADD3 + IMAD.X per row)
//   fma_pipe_load / fma_pipe_store   base[idx] with a 32-bit element index: IMAD.WIDE.U32 + IMAD.  The element size is
//                   passed as a 64-bit value `four` the compiler cannot see through (callers derive it from loaded data:
//                   4 + (x >> 40) with x < 2^40): with an immediate 4 ptxas emits LEA + LEA.HI.X, with mad.wide.u32 and a
//                   64-bit addend it splits the product from an IADD3 + IMAD.X add -- ALU pipe both times (sm_100a,
//                   CUDA 12.9; SASS checked with cuobjdump).
__device__ __forceinline__ uint32_t fma_pipe_mad(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint64_t fma_pipe_addr(const void* base, uint32_t idx, uint64_t elem_bytes) {
    uint64_t a;
    asm("{\n\t.reg .u64 t;\n\tcvt.u64.u32 t, %1;\n\tmad.lo.u64 %0, t, %2, %3;\n\t}"
        : "=l"(a) : "r"(idx), "l"(elem_bytes), "l"(base));
    return a;
}
__device__ __forceinline__ float fma_pipe_load(const float* base, uint32_t idx, uint64_t four) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(fma_pipe_addr(base, idx, four)));
    return v;
}
__device__ __forceinline__ void fma_pipe_store(float* base, uint32_t idx, uint64_t four, float v) {
    asm volatile("st.global.f32 [%0], %1;" ::"l"(fma_pipe_addr(base, idx, four)), "f"(v) : "memory");
}

// order-preserving float <-> uint mapping for atomicMax on floats of either sign
__device__ __forceinline__ uint32_t float_to_ordered(float x) {
    const uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t k) {
    const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}
#endif

}  // namespace hpss
