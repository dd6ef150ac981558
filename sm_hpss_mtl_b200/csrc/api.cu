// C-ABI layer of libhpss_b200: context / plan caches / batch layouts / fused and host entry
// points.  Declarations and the reference call each entry replaces: include/hpss_b200.h.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>

#include "common.cuh"

namespace hpss {

static thread_local std::string t_error;
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    t_error = buf;
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return HPSS_ERR_CUDA;
}

// development knobs: the environment is read once, never on a launch path
const Knobs& knobs() {
    static const Knobs k = [] {
        auto geti = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
        auto has = [](const char* name) { return getenv(name) != nullptr ? 1 : 0; };
        Knobs v;
        v.no_sweep = has("HPSS_NO_SWEEP");
        v.host_chunks = std::max(0, geti("HPSS_HOST_CHUNKS", 0));
        v.no_uniform_moments = has("HPSS_NO_UNIFORM_MOMENTS");
        v.mom_ctas = geti("HPSS_MOM_CTAS", 16) > 0 ? geti("HPSS_MOM_CTAS", 16) : 16;
        v.no_fast_stft = has("HPSS_NO_FAST_STFT");
        v.no_uniform_stft = has("HPSS_NO_UNIFORM_STFT");
        v.k1_real = geti("HPSS_K1_REAL", 1);
        v.sweep1 = has("HPSS_SWEEP1");
        v.sweep_u = geti("HPSS_SWEEP_U", 4);
        v.dct_grid_mult = geti("HPSS_DCT_GRID_MULT", 8);
        v.no_dense_median = has("HPSS_NO_DENSE_MEDIAN");
        return v;
    }();
    return k;
}

// ---- host-side tables --------------------------------------------------------------
static double hz_to_mel(double f) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m) {
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp;
    const double logstep = log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}
// numpy.linspace(start, stop, num, endpoint=True)
static void linspace(double start, double stop, int num, std::vector<double>& y) {
    y.resize(num);
    if (num == 1) { y[0] = start; return; }
    const double step = (stop - start) / (double)(num - 1);
    for (int i = 0; i < num; ++i) y[i] = (double)i * step + start;
    y[num - 1] = stop;
}

// librosa.filters.mel(sr, n_fft, n_mels, fmin=0, fmax=sr/2, htk=False, norm='slaney', dtype=float32)
int build_mel(int sr, int n_fft, int n_mels, float* out) {
    if (sr <= 0 || n_fft < 2 || n_mels < 1) {
        set_error("mel: invalid sr=%d n_fft=%d n_mels=%d", sr, n_fft, n_mels);
        return HPSS_ERR_INVALID;
    }
    const int nfreq = 1 + n_fft / 2;
    std::vector<double> fftfreqs, mels, mel_f(n_mels + 2);
    linspace(0.0, (double)sr / 2, nfreq, fftfreqs);
    linspace(hz_to_mel(0.0), hz_to_mel((double)sr / 2), n_mels + 2, mels);
    for (int i = 0; i < n_mels + 2; ++i) mel_f[i] = mel_to_hz(mels[i]);
    for (int i = 0; i < n_mels; ++i) {
        const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
        const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
        for (int f = 0; f < nfreq; ++f) {
            const double lower = -(mel_f[i] - fftfreqs[f]) / fd0;
            const double upper = (mel_f[i + 2] - fftfreqs[f]) / fd1;
            const float w = (float)std::max(0.0, std::min(lower, upper));   // stored as float32 ...
            out[(size_t)i * nfreq + f] = (float)((double)w * enorm);          // ... then *= enorm (f64 -> f32)
        }
    }
    return HPSS_OK;
}

// scipy.signal.get_window('hann', win, fftbins=True), librosa.util.pad_center to n_fft
void build_window(int n_fft, int win, float* out) {
    const int lpad = (n_fft - win) / 2;
    for (int i = 0; i < n_fft; ++i) out[i] = 0.f;
    for (int n = 0; n < win; ++n) out[lpad + n] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * (double)n / (double)win));
}

static int factor_radices(int n2, int* radix, int* n_pass) {
    int n = n2, np = 0;
    const int odd[2] = {5, 3};
    for (int p : odd)
        while (n % p == 0) { if (np >= kMaxRadixPasses) return -1; radix[np++] = p; n /= p; }
    while (n % 4 == 0) { if (np >= kMaxRadixPasses) return -1; radix[np++] = 4; n /= 4; }
    while (n % 2 == 0) { if (np >= kMaxRadixPasses) return -1; radix[np++] = 2; n /= 2; }
    if (n != 1 || np == 0) return -1;
    *n_pass = np;
    return 0;
}

int get_fft_plan(hpss_ctx* ctx, int n_fft, int win, FftPlan** out) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    auto key = std::make_pair(n_fft, win);
    auto it = ctx->fft_plans.find(key);
    if (it != ctx->fft_plans.end()) { *out = it->second; return HPSS_OK; }
    if (n_fft < 4 || (n_fft & 1)) {
        set_error("n_fft=%d must be even and >= 4", n_fft);
        return HPSS_ERR_UNSUPPORTED;
    }
    if (win < 1 || win > n_fft) {
        set_error("win_length=%d must be in [1, n_fft=%d] (librosa.util.pad_center)", win, n_fft);
        return HPSS_ERR_INVALID;
    }
    FftPlan* p = new FftPlan();
    p->n_fft = n_fft; p->win = win; p->n2 = n_fft / 2;
    if (factor_radices(p->n2, p->radix, &p->n_pass)) {
        delete p;
        set_error("n_fft=%d: n_fft/2 must factor into 2, 3 and 5", n_fft);
        return HPSS_ERR_UNSUPPORTED;
    }
    std::vector<float> w(n_fft);
    build_window(n_fft, win, w.data());
    std::vector<float2> th(p->n2), tf(p->n2 + 1);
    for (int k = 0; k < p->n2; ++k) {
        const double a = -2.0 * M_PI * (double)k / (double)p->n2;
        th[k] = make_float2((float)cos(a), (float)sin(a));
    }
    for (int k = 0; k <= p->n2; ++k) {
        const double a = -2.0 * M_PI * (double)k / (double)n_fft;
        tf[k] = make_float2((float)cos(a), (float)sin(a));
    }
    HPSS_CUDA(cudaMalloc(&p->d_window, sizeof(float) * n_fft));
    HPSS_CUDA(cudaMalloc(&p->d_window_half, sizeof(float) * n_fft));
    HPSS_CUDA(cudaMalloc(&p->d_tw_half, sizeof(float2) * p->n2));
    HPSS_CUDA(cudaMalloc(&p->d_tw_full, sizeof(float2) * (p->n2 + 1)));
    HPSS_CUDA(cudaMemcpy(p->d_window, w.data(), sizeof(float) * n_fft, cudaMemcpyHostToDevice));
    for (auto& v : w) v *= 0.5f;       // exact; the specialised kernel keeps Z/2 so that the unpack needs no halving
    HPSS_CUDA(cudaMemcpy(p->d_window_half, w.data(), sizeof(float) * n_fft, cudaMemcpyHostToDevice));
    HPSS_CUDA(cudaMemcpy(p->d_tw_half, th.data(), sizeof(float2) * p->n2, cudaMemcpyHostToDevice));
    HPSS_CUDA(cudaMemcpy(p->d_tw_full, tf.data(), sizeof(float2) * (p->n2 + 1), cudaMemcpyHostToDevice));
    int na = 0, nb = 0;
    if (stft_fast_split(n_fft, &na, &nb)) {
        std::vector<float2> wbq((size_t)p->n2), tkb((size_t)p->n2);
        for (int b = 0; b < nb; ++b)
            for (int q = 0; q < na; ++q) {
                const int n = nb * q + b;
                wbq[(size_t)b * na + q] = make_float2(w[2 * n], w[2 * n + 1]);         // w is already halved here
            }
        for (int k1 = 0; k1 < na; ++k1)
            for (int b = 0; b < nb; ++b) tkb[(size_t)k1 * nb + b] = th[(b * k1) % p->n2];
        HPSS_CUDA(cudaMalloc(&p->d_win_bq, sizeof(float2) * p->n2));
        HPSS_CUDA(cudaMalloc(&p->d_tw_kb, sizeof(float2) * p->n2));
        HPSS_CUDA(cudaMemcpy(p->d_win_bq, wbq.data(), sizeof(float2) * p->n2, cudaMemcpyHostToDevice));
        HPSS_CUDA(cudaMemcpy(p->d_tw_kb, tkb.data(), sizeof(float2) * p->n2, cudaMemcpyHostToDevice));
    }
    if (n_fft == 400) {
        std::vector<float> wr(400);
        std::vector<float2> tr(11 * 20);
        for (int b = 0; b < 20; ++b)
            for (int q = 0; q < 20; ++q) wr[b * 20 + q] = 2.0f * w[20 * q + b];       // w was halved above (exactly)
        for (int k1 = 0; k1 <= 10; ++k1)
            for (int b = 0; b < 20; ++b) tr[k1 * 20 + b] = tf[b * k1];
        HPSS_CUDA(cudaMalloc(&p->d_win_r400, sizeof(float) * wr.size()));
        HPSS_CUDA(cudaMalloc(&p->d_tw_r400, sizeof(float2) * tr.size()));
        HPSS_CUDA(cudaMemcpy(p->d_win_r400, wr.data(), sizeof(float) * wr.size(), cudaMemcpyHostToDevice));
        HPSS_CUDA(cudaMemcpy(p->d_tw_r400, tr.data(), sizeof(float2) * tr.size(), cudaMemcpyHostToDevice));
    }
    ctx->fft_plans[key] = p;
    *out = p;
    return HPSS_OK;
}

int get_mel_plan(hpss_ctx* ctx, int sr, int n_fft, int n_mels, MelPlan** out) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    auto key = std::make_tuple(sr, n_fft, n_mels);
    auto it = ctx->mel_plans.find(key);
    if (it != ctx->mel_plans.end()) { *out = it->second; return HPSS_OK; }
    const int rows = 1 + n_fft / 2;
    std::vector<float> w((size_t)n_mels * rows);
    int rc = build_mel(sr, n_fft, n_mels, w.data());
    if (rc) return rc;
    std::vector<int2> band(n_mels);
    for (int m = 0; m < n_mels; ++m) {
        int first = rows, last = 0;
        for (int f = 0; f < rows; ++f)
            if (w[(size_t)m * rows + f] != 0.f) { first = std::min(first, f); last = f + 1; }
        if (first >= last) first = last = 0;
        band[m] = make_int2(first, last);
    }
    MelPlan* p = new MelPlan();
    p->sr = sr; p->n_fft = n_fft; p->n_mels = n_mels; p->rows = rows;
    // sweep table: filters ordered, at most two overlapping at any frequency row
    std::vector<int4> sweep(rows);
    bool ok = true;
    {
        int pa = -1, pb = -1;
        for (int m = 0; m < n_mels && ok; ++m) {
            if (band[m].y <= band[m].x) continue;                    // empty filter
            if (band[m].x < pa || band[m].y < pb) ok = false;        // bands must be ordered
            pa = band[m].x; pb = band[m].y;
        }
        for (int f = 0; f < rows && ok; ++f) {
            int first = n_mels;
            for (int m = 0; m < n_mels; ++m) if (band[m].y > f) { first = m; break; }
            for (int m = 0; m < n_mels; ++m)
                if (w[(size_t)m * rows + f] != 0.f && (m < first || m > first + 1)) ok = false;
            const float wa = first < n_mels ? w[(size_t)first * rows + f] : 0.f;
            const float wb = first + 1 < n_mels ? w[(size_t)(first + 1) * rows + f] : 0.f;
            int ia, ib;
            memcpy(&ia, &wa, 4); memcpy(&ib, &wb, 4);
            sweep[f] = make_int4(first, ia, ib, 0);
        }
    }
    p->sweepable = ok;
    if (ok) {
        HPSS_CUDA(cudaMalloc(&p->d_sweep, sizeof(int4) * rows));
        HPSS_CUDA(cudaMemcpy(p->d_sweep, sweep.data(), sizeof(int4) * rows, cudaMemcpyHostToDevice));
        // emission counts (4 bits per row; padded with zeros to a multiple of 64 rows) and weights per row
        const int rows_pad = (rows + 63) / 64 * 64;
        std::vector<uint32_t> emit4(rows_pad / 8, 0u);
        std::vector<float2> sw(rows_pad + 1, make_float2(0.f, 0.f));   // the sweep reads one row ahead
        bool walk = true;
        int prev = 0;
        for (int f = 0; f < rows; ++f) {
            const int cnt = sweep[f].x - prev;
            prev = sweep[f].x;
            if (cnt < 0 || cnt > 15) { walk = false; break; }
            emit4[f / 8] |= (uint32_t)cnt << (4 * (f % 8));
            float wa, wb;
            memcpy(&wa, &sweep[f].y, 4); memcpy(&wb, &sweep[f].z, 4);
            sw[f] = make_float2(wa, wb);
        }
        p->walkable = walk;
        if (walk) {
            HPSS_CUDA(cudaMalloc(&p->d_emit4, sizeof(uint32_t) * emit4.size()));
            HPSS_CUDA(cudaMalloc(&p->d_sweep_w, sizeof(float2) * sw.size()));
            HPSS_CUDA(cudaMemcpy(p->d_emit4, emit4.data(), sizeof(uint32_t) * emit4.size(), cudaMemcpyHostToDevice));
            HPSS_CUDA(cudaMemcpy(p->d_sweep_w, sw.data(), sizeof(float2) * sw.size(), cudaMemcpyHostToDevice));
        }
    }
    HPSS_CUDA(cudaMalloc(&p->d_w, sizeof(float) * w.size()));
    HPSS_CUDA(cudaMalloc(&p->d_band, sizeof(int2) * n_mels));
    HPSS_CUDA(cudaMemcpy(p->d_w, w.data(), sizeof(float) * w.size(), cudaMemcpyHostToDevice));
    HPSS_CUDA(cudaMemcpy(p->d_band, band.data(), sizeof(int2) * n_mels, cudaMemcpyHostToDevice));
    ctx->mel_plans[key] = p;
    *out = p;
    return HPSS_OK;
}

int ensure_workspace(hpss_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->ws_bytes) return HPSS_OK;
    if (ctx->ws) {
        HPSS_CUDA(cudaDeviceSynchronize());   // rare: only when the workspace has to grow
        HPSS_CUDA(cudaFree(ctx->ws));
        ctx->ws = nullptr;
        ctx->ws_bytes = 0;
    }
    const size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&ctx->ws, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaMalloc(&ctx->ws, bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("out of device memory: workspace of %zu bytes", bytes);
            return HPSS_ERR_NOMEM;
        }
        ctx->ws_bytes = bytes;
        return HPSS_OK;
    }
    ctx->ws_bytes = want;
    return HPSS_OK;
}

int ensure_stft_tiles(hpss_batch* b, int tt) {
    if (b->stft_tt == tt && (b->d_stft_tiles || b->n_stft_tiles == 0)) return HPSS_OK;
    std::vector<int2> tiles;
    for (int c = 0; c < b->n_clips; ++c) {
        const int64_t T = b->frame_off[c + 1] - b->frame_off[c];
        for (int64_t t0 = 0; t0 < T; t0 += tt) tiles.push_back(make_int2(c, (int)t0));
    }
    if (b->d_stft_tiles) { HPSS_CUDA(cudaFree(b->d_stft_tiles)); b->d_stft_tiles = nullptr; }
    b->n_stft_tiles = (int)tiles.size();
    b->stft_tt = tt;
    if (!tiles.empty()) {
        HPSS_CUDA(cudaMalloc(&b->d_stft_tiles, sizeof(int2) * tiles.size()));
        HPSS_CUDA(cudaMemcpy(b->d_stft_tiles, tiles.data(), sizeof(int2) * tiles.size(), cudaMemcpyHostToDevice));
    }
    return HPSS_OK;
}

static int make_batch(hpss_ctx* ctx, const std::vector<int64_t>& samples, const std::vector<int64_t>& frames,
                      bool has_samples, int n_fft, int hop, hpss_batch** out) {
    const int n = (int)frames.size();
    hpss_batch* b = new hpss_batch();
    b->ctx = ctx; b->n_clips = n; b->n_fft = n_fft; b->hop = hop; b->has_samples = has_samples;
    b->sample_off.assign(n + 1, 0);
    b->frame_off.assign(n + 1, 0);
    for (int c = 0; c < n; ++c) {
        b->sample_off[c + 1] = b->sample_off[c] + (has_samples ? samples[c] : 0);
        b->frame_off[c + 1] = b->frame_off[c] + frames[c];
        b->max_frames = std::max(b->max_frames, frames[c]);
    }
    b->uniform_frames = (n > 0) ? frames[0] : 0;
    for (int c = 1; c < n; ++c)
        if (frames[c] != frames[0]) { b->uniform_frames = 0; break; }
    b->uniform_samples = (n > 0 && has_samples) ? samples[0] : 0;
    for (int c = 1; c < n && b->uniform_samples; ++c)
        if (samples[c] != samples[0]) b->uniform_samples = 0;
    // clip of the first frame of every 32-frame block (empty clips are skipped)
    const int64_t total = b->frame_off[n];
    std::vector<int32_t> block_clip((size_t)((total + 31) / 32) + 1, 0);
    {
        int c = 0;
        for (size_t blk = 0; blk + 1 < block_clip.size() + 0; ++blk) {
            const int64_t g = (int64_t)blk * 32;
            while (c < n && b->frame_off[c + 1] <= g) ++c;
            block_clip[blk] = c < n ? c : (n > 0 ? n - 1 : 0);
        }
    }
    cudaError_t e = cudaMalloc(&b->d_frame_off, sizeof(int64_t) * (n + 1));
    if (e == cudaSuccess) e = cudaMalloc(&b->d_block_clip, sizeof(int32_t) * block_clip.size());
    if (e == cudaSuccess)
        e = cudaMemcpy(b->d_block_clip, block_clip.data(), sizeof(int32_t) * block_clip.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&b->d_sample_off, sizeof(int64_t) * (n + 1));
    if (e == cudaSuccess)
        e = cudaMemcpy(b->d_frame_off, b->frame_off.data(), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
        e = cudaMemcpy(b->d_sample_off, b->sample_off.data(), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        if (b->d_frame_off) cudaFree(b->d_frame_off);
        if (b->d_sample_off) cudaFree(b->d_sample_off);
        if (b->d_block_clip) cudaFree(b->d_block_clip);
        delete b;
        return cuda_fail(e, "batch offsets");
    }
    *out = b;
    return HPSS_OK;
}

static int feature_streams(int feature) { return feature >= HPSS_FEAT_HARMPERC ? 2 : 1; }
static bool feature_is_mel(int feature) {
    return feature == HPSS_FEAT_MELSPEC || feature == HPSS_FEAT_LOGMELSPEC || feature == HPSS_FEAT_MEL_HARMPERC ||
           feature == HPSS_FEAT_LOGMEL_HARMPERC;
}
static bool feature_is_log(int feature) {
    return feature == HPSS_FEAT_LOGSPEC || feature == HPSS_FEAT_LOGMELSPEC || feature == HPSS_FEAT_LOG_HARMPERC ||
           feature == HPSS_FEAT_LOGMEL_HARMPERC;
}

static int check_params(const hpss_params* p) {
    if (!p) { set_error("params is NULL"); return HPSS_ERR_INVALID; }
    if (p->feature < HPSS_FEAT_SPEC || p->feature > HPSS_FEAT_LOGMEL_HARMPERC) {
        set_error("unknown feature id %d", p->feature);
        return HPSS_ERR_INVALID;
    }
    if (feature_is_mel(p->feature) && (p->n_mels < 1 || p->mel_sr < 1)) {
        set_error("feature %d needs n_mels >= 1 and mel_sr >= 1 (got %d, %d)", p->feature, p->n_mels, p->mel_sr);
        return HPSS_ERR_INVALID;
    }
    if (feature_streams(p->feature) == 2 && (p->l_harm < 1 || p->l_perc < 1)) {
        set_error("l_harm=%d / l_perc=%d must be >= 1", p->l_harm, p->l_perc);
        return HPSS_ERR_INVALID;
    }
    if (feature_is_log(p->feature) && !(p->amin > 0.f)) {
        set_error("amin must be strictly positive (librosa.power_to_db)");
        return HPSS_ERR_INVALID;
    }
    return HPSS_OK;
}

// optional sink of the fused "top_db clip + feature moments" tail
struct MomentSink {
    const int32_t* d_class = nullptr;
    int n_classes = 0;
    double *sum = nullptr, *sumsq = nullptr, *count = nullptr, *nonfinite = nullptr;
};

// spectrogram -> features (everything after the STFT); S may alias nothing in the workspace
static int features_from_spec(hpss_ctx* ctx, const hpss_batch* b, const float* S, int rows, const hpss_params* p,
                              float* harm, float* perc, uint32_t* clip_max, float* out, cudaStream_t st,
                              const MomentSink* ms = nullptr) {
    const int ns = feature_streams(p->feature);
    const bool is_mel = feature_is_mel(p->feature), is_log = feature_is_log(p->feature);
    MelPlan* mp = nullptr;
    int rc;
    if (is_mel) {
        rc = get_mel_plan(ctx, p->mel_sr, 2 * (rows - 1), p->n_mels, &mp);
        if (rc) return rc;
    }
    const bool clip = is_log && p->top_db >= 0.f;
    if (ns == 2) {
        rc = launch_median(ctx, b, S, rows, p->l_harm, true, harm, st);
        if (rc) return rc;
        rc = launch_median(ctx, b, S, rows, p->l_perc, false, perc, st);
        if (rc) return rc;
    }
    {
        const int pre_square = (ns == 1 && is_mel) ? 1 : 0;   // melspectrogram(y=..) uses |X|^2
        rc = launch_mask_mel(ctx, b, S, ns == 2 ? harm : nullptr, ns == 2 ? perc : nullptr, rows,
                             mp ? mp->d_w : nullptr, mp ? mp->d_band : nullptr,
                             (mp && mp->sweepable && !knobs().no_sweep) ? mp->d_sweep : nullptr,
                             (mp && mp->walkable) ? mp->d_emit4 : nullptr, (mp && mp->walkable) ? mp->d_sweep_w : nullptr,
                             mp ? mp->n_mels : 0, pre_square,
                             is_log ? 1 : 0, p->amin, out, clip ? clip_max : nullptr, st);
        if (rc) return rc;
    }
    const int rps = is_mel ? p->n_mels : rows;
    if (ms) {   // one pass: clip (if any) + moments
        if (clip)
            return launch_topdb_moments(ctx, b, out, rps, ns, clip_max, p->top_db, ms->d_class, ms->n_classes, ms->sum,
                                        ms->sumsq, ms->count, ms->nonfinite, st);
        return launch_moments(ctx, b, out, rps * ns, ms->d_class, ms->n_classes, ms->sum, ms->sumsq, ms->count,
                              ms->nonfinite, st);
    }
    if (clip) {
        rc = launch_topdb(ctx, b, out, rps, ns, clip_max, p->top_db, st);
        if (rc) return rc;
    }
    return HPSS_OK;
}

struct Workspace {
    float* S; float* harm; float* perc; uint32_t* clip_max;
};
static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
static int carve_workspace(hpss_ctx* ctx, const hpss_batch* b, int rows, bool need_S, bool need_hp, size_t extra,
                           Workspace* w, void** extra_ptr) {
    const size_t n = (size_t)rows * (size_t)b->frame_off[b->n_clips];
    const size_t sz = align256(n * sizeof(float));
    const size_t cm = align256(sizeof(uint32_t) * 2 * (size_t)std::max(1, b->n_clips));
    const size_t total = (need_S ? sz : 0) + (need_hp ? 2 * sz : 0) + cm + align256(extra);
    int rc = ensure_workspace(ctx, total);
    if (rc) return rc;
    char* p = (char*)ctx->ws;
    w->S = nullptr; w->harm = nullptr; w->perc = nullptr;
    if (need_S) { w->S = (float*)p; p += sz; }
    if (need_hp) { w->harm = (float*)p; p += sz; w->perc = (float*)p; p += sz; }
    w->clip_max = (uint32_t*)p; p += cm;
    if (extra_ptr) *extra_ptr = p;
    return HPSS_OK;
}

static int featuregram_device(hpss_ctx* ctx, hpss_batch* b, const float* wave, const hpss_params* p, float* out,
                              cudaStream_t st, const MomentSink* ms = nullptr) {
    int rc = check_params(p);
    if (rc) return rc;
    if (!b->has_samples) { set_error("batch was built from frame counts; waveform entry needs sample lengths"); return HPSS_ERR_INVALID; }
    if (b->n_fft != p->n_fft || b->hop != p->hop_length) {
        set_error("batch was laid out for n_fft=%d hop=%d, params say %d / %d", b->n_fft, b->hop, p->n_fft, p->hop_length);
        return HPSS_ERR_INVALID;
    }
    FftPlan* plan = nullptr;
    rc = get_fft_plan(ctx, p->n_fft, p->win_length, &plan);
    if (rc) return rc;
    const int rows = p->n_fft / 2 + 1;
    if (p->feature == HPSS_FEAT_SPEC) {
        rc = launch_stft(ctx, b, wave, plan, p->hop_length, 0, out, nullptr, st);
        if (rc || !ms) return rc;
        return launch_moments(ctx, b, out, rows, ms->d_class, ms->n_classes, ms->sum, ms->sumsq, ms->count,
                              ms->nonfinite, st);
    }
    Workspace w;
    rc = carve_workspace(ctx, b, rows, true, feature_streams(p->feature) == 2, 0, &w, nullptr);
    if (rc) return rc;
    rc = launch_stft(ctx, b, wave, plan, p->hop_length, 0, w.S, nullptr, st);
    if (rc) return rc;
    return features_from_spec(ctx, b, w.S, rows, p, w.harm, w.perc, w.clip_max, out, st, ms);
}

}  // namespace hpss

using namespace hpss;

// =====================================================================================
extern "C" {

const char* hpss_version(void) { return "hpss_b200 0.1 (sm_100a)"; }
const char* hpss_last_error(void) { return t_error.c_str(); }
uint64_t hpss_launch_count(void) { return g_launches.load(); }

int hpss_ctx_create(int device, hpss_ctx** out) {
    if (!out) { set_error("ctx out pointer is NULL"); return HPSS_ERR_INVALID; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); this library has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return HPSS_ERR_CUDA;
    }
    if (device < 0 || device >= n) { set_error("device %d out of range [0,%d)", device, n); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(device));
    hpss_ctx* ctx = new hpss_ctx();
    ctx->device = device;
    cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (ctx->sm_count <= 0) ctx->sm_count = kSMs;
    cudaError_t e2 = cudaEventCreateWithFlags(&ctx->ws_done, cudaEventDisableTiming);
    if (e2 == cudaSuccess) e2 = cudaMalloc(&ctx->d_flags, 4 * sizeof(uint32_t));
    if (e2 == cudaSuccess) e2 = cudaMemset(ctx->d_flags, 0, 4 * sizeof(uint32_t));
    if (e2 == cudaSuccess) e2 = cudaHostAlloc(&ctx->h_flags, 4 * sizeof(uint32_t), cudaHostAllocDefault);
    if (e2 != cudaSuccess) {
        if (ctx->ws_done) cudaEventDestroy(ctx->ws_done);
        if (ctx->d_flags) cudaFree(ctx->d_flags);
        delete ctx;
        return cuda_fail(e2, "context status word");
    }
    *out = ctx;
    return HPSS_OK;
}

int hpss_ctx_destroy(hpss_ctx* ctx) {
    if (!ctx) return HPSS_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto& kv : ctx->fft_plans) {
        cudaFree(kv.second->d_window); cudaFree(kv.second->d_window_half); cudaFree(kv.second->d_win_bq); cudaFree(kv.second->d_win_r400); cudaFree(kv.second->d_tw_r400); cudaFree(kv.second->d_tw_kb); cudaFree(kv.second->d_tw_half); cudaFree(kv.second->d_tw_full);
        delete kv.second;
    }
    for (auto& kv : ctx->dct_plans) cudaFree(kv.second);
    for (auto& kv : ctx->mel_plans) { cudaFree(kv.second->d_w); cudaFree(kv.second->d_band); if (kv.second->d_sweep) cudaFree(kv.second->d_sweep); if (kv.second->d_emit4) cudaFree(kv.second->d_emit4); if (kv.second->d_sweep_w) cudaFree(kv.second->d_sweep_w); delete kv.second; }
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->prep_ws) cudaFree(ctx->prep_ws);
    if (ctx->d_flags) cudaFree(ctx->d_flags);
    if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
    if (ctx->ws_done) cudaEventDestroy(ctx->ws_done);
    if (ctx->band_scratch) cudaFree(ctx->band_scratch);
    for (int i = 0; i < 2; ++i) {
        if (ctx->pipe_dev[i]) cudaFree(ctx->pipe_dev[i]);
        if (ctx->ev_h2d[i]) cudaEventDestroy(ctx->ev_h2d[i]);
        if (ctx->ev_comp[i]) cudaEventDestroy(ctx->ev_comp[i]);
        if (ctx->ev_d2h[i]) cudaEventDestroy(ctx->ev_d2h[i]);
    }
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_comp) cudaStreamDestroy(ctx->s_comp);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    delete ctx;
    return HPSS_OK;
}

int hpss_ctx_device(const hpss_ctx* ctx) { return ctx ? ctx->device : -1; }
uint64_t hpss_ctx_workspace_bytes(const hpss_ctx* ctx) { return ctx ? ctx->ws_bytes + 2 * ctx->pipe_bytes + ctx->prep_ws_bytes : 0; }

int hpss_host_alloc(void** ptr, uint64_t bytes) {
    if (!ptr) { set_error("ptr is NULL"); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return HPSS_OK;
}
int hpss_host_free(void* ptr) {
    if (ptr) HPSS_CUDA(cudaFreeHost(ptr));
    return HPSS_OK;
}

int hpss_batch_from_samples(hpss_ctx* ctx, const int64_t* clip_len, int32_t n_clips, int32_t n_fft, int32_t hop,
                            hpss_batch** out) {
    if (!ctx || !out || n_clips < 0 || (n_clips > 0 && !clip_len)) { set_error("batch_from_samples: bad arguments"); return HPSS_ERR_INVALID; }
    if (n_fft < 1 || hop < 1) { set_error("n_fft=%d and hop_length=%d must be positive", n_fft, hop); return HPSS_ERR_INVALID; }
    std::vector<int64_t> s(n_clips), f(n_clips);
    for (int c = 0; c < n_clips; ++c) {
        if (clip_len[c] < n_fft) {
            set_error("n_fft=%d is too large for input signal of length=%lld (clip %d)", n_fft, (long long)clip_len[c], c);
            return HPSS_ERR_SHORT_SIGNAL;
        }
        s[c] = clip_len[c];
        f[c] = 1 + (clip_len[c] - n_fft) / hop;
    }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    return make_batch(ctx, s, f, true, n_fft, hop, out);
}

int hpss_batch_from_frames(hpss_ctx* ctx, const int64_t* clip_frames, int32_t n_clips, hpss_batch** out) {
    if (!ctx || !out || n_clips < 0 || (n_clips > 0 && !clip_frames)) { set_error("batch_from_frames: bad arguments"); return HPSS_ERR_INVALID; }
    std::vector<int64_t> s, f(n_clips);
    for (int c = 0; c < n_clips; ++c) {
        if (clip_frames[c] < 0 || clip_frames[c] > 0x7fffffff) { set_error("clip %d: frame count %lld out of range", c, (long long)clip_frames[c]); return HPSS_ERR_INVALID; }
        f[c] = clip_frames[c];
    }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    return make_batch(ctx, s, f, false, 0, 0, out);
}

int hpss_batch_destroy(hpss_batch* b) {
    if (!b) return HPSS_OK;
    cudaSetDevice(b->ctx->device);
    cudaDeviceSynchronize();
    if (b->d_frame_off) cudaFree(b->d_frame_off);
    if (b->d_sample_off) cudaFree(b->d_sample_off);
    if (b->d_block_clip) cudaFree(b->d_block_clip);
    if (b->d_stft_tiles) cudaFree(b->d_stft_tiles);
    if (b->d_clip_class) cudaFree(b->d_clip_class);
    for (auto& kv : b->time_tiles) if (kv.second.first) cudaFree(kv.second.first);
    for (auto& kv : b->patch_offs) if (kv.second.second) cudaFree(kv.second.second);
    if (b->host_pipe) hpss_pipeline_destroy(b->host_pipe);
    delete b;
    return HPSS_OK;
}
int32_t hpss_batch_n_clips(const hpss_batch* b) { return b ? b->n_clips : 0; }
int64_t hpss_batch_total_frames(const hpss_batch* b) { return b ? b->frame_off[b->n_clips] : 0; }
int64_t hpss_batch_total_samples(const hpss_batch* b) { return b ? b->sample_off[b->n_clips] : 0; }
int hpss_batch_frame_offsets(const hpss_batch* b, int64_t* o) {
    if (!b || !o) { set_error("frame_offsets: NULL argument"); return HPSS_ERR_INVALID; }
    memcpy(o, b->frame_off.data(), sizeof(int64_t) * (b->n_clips + 1));
    return HPSS_OK;
}
int hpss_batch_sample_offsets(const hpss_batch* b, int64_t* o) {
    if (!b || !o) { set_error("sample_offsets: NULL argument"); return HPSS_ERR_INVALID; }
    memcpy(o, b->sample_off.data(), sizeof(int64_t) * (b->n_clips + 1));
    return HPSS_OK;
}

int hpss_mel_filterbank(int32_t sr, int32_t n_fft, int32_t n_mels, float* out) {
    if (!out) { set_error("out is NULL"); return HPSS_ERR_INVALID; }
    return build_mel(sr, n_fft, n_mels, out);
}
int hpss_stft_window(int32_t n_fft, int32_t win, float* out) {
    if (!out || win < 1 || win > n_fft) { set_error("stft_window: need 1 <= win_length <= n_fft"); return HPSS_ERR_INVALID; }
    build_window(n_fft, win, out);
    return HPSS_OK;
}

int hpss_stft_mag(hpss_ctx* ctx, const hpss_batch* batch, const float* wave, int32_t n_fft, int32_t win,
                  int32_t hop, int32_t power, float* S, float* cplx, void* stream) {
    if (!ctx || !batch || !wave || !S) { set_error("stft_mag: NULL argument"); return HPSS_ERR_INVALID; }
    if (!batch->has_samples || batch->n_fft != n_fft || batch->hop != hop) {
        set_error("stft_mag: batch layout (n_fft=%d hop=%d) does not match call (n_fft=%d hop=%d)", batch->n_fft,
                  batch->hop, n_fft, hop);
        return HPSS_ERR_INVALID;
    }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    FftPlan* plan = nullptr;
    int rc = get_fft_plan(ctx, n_fft, win, &plan);
    if (rc) return rc;
    return launch_stft(ctx, const_cast<hpss_batch*>(batch), wave, plan, hop, power, S, cplx, (cudaStream_t)stream);
}

int hpss_median_time(hpss_ctx* ctx, const hpss_batch* batch, const float* S, int32_t rows, int32_t k, float* out,
                     void* stream) {
    if (!ctx || !batch || !S || !out || S == out) { set_error("median_time: NULL or aliased argument"); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    return launch_median(ctx, batch, S, rows, k, true, out, (cudaStream_t)stream);
}
int hpss_median_freq(hpss_ctx* ctx, const hpss_batch* batch, const float* S, int32_t rows, int32_t k, float* out,
                     void* stream) {
    if (!ctx || !batch || !S || !out || S == out) { set_error("median_freq: NULL or aliased argument"); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    return launch_median(ctx, batch, S, rows, k, false, out, (cudaStream_t)stream);
}

int hpss_mask_mel_log(hpss_ctx* ctx, const hpss_batch* batch, const float* S, const float* harm, const float* perc,
                      int32_t rows, const float* mel, int32_t n_mels, int32_t pre_square, int32_t log_power,
                      float amin, float* out, uint32_t* clip_max, void* stream) {
    if (!ctx || !batch || !S || !out) { set_error("mask_mel_log: NULL argument"); return HPSS_ERR_INVALID; }
    if ((harm == nullptr) != (perc == nullptr)) { set_error("mask_mel_log: harm and perc must both be given or both NULL"); return HPSS_ERR_INVALID; }
    if (rows < 1 || (mel && n_mels < 1)) { set_error("mask_mel_log: rows=%d n_mels=%d", rows, n_mels); return HPSS_ERR_INVALID; }
    if (log_power && !(amin > 0.f)) { set_error("amin must be strictly positive (librosa.power_to_db)"); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    int2* band = nullptr;
    ScratchLease lease;                          // band_scratch is context scratch
    {
        const int rc = lease.begin(ctx, st);
        if (rc) return rc;
    }
    if (mel) {
        if (ctx->band_scratch_n < n_mels) {
            if (ctx->band_scratch) { HPSS_CUDA(cudaDeviceSynchronize()); HPSS_CUDA(cudaFree(ctx->band_scratch)); ctx->band_scratch = nullptr; }
            HPSS_CUDA(cudaMalloc(&ctx->band_scratch, sizeof(int2) * n_mels));
            ctx->band_scratch_n = n_mels;
        }
        band = ctx->band_scratch;
        int rc = launch_mel_bands(mel, n_mels, rows, band, st);
        if (rc) return rc;
    }
    return launch_mask_mel(ctx, batch, S, harm, perc, rows, mel, band, nullptr, nullptr, nullptr, n_mels, pre_square,
                           log_power, amin, out, clip_max, st);
}

int hpss_mask_mel_log_sr(hpss_ctx* ctx, const hpss_batch* batch, const float* S, const float* harm, const float* perc,
                         int32_t rows, int32_t mel_sr, int32_t n_mels, int32_t pre_square, int32_t log_power,
                         float amin, float* out, uint32_t* clip_max, void* stream) {
    if (!ctx || !batch || !S || !out) { set_error("mask_mel_log_sr: NULL argument"); return HPSS_ERR_INVALID; }
    if ((harm == nullptr) != (perc == nullptr)) { set_error("mask_mel_log_sr: harm and perc must both be given or both NULL"); return HPSS_ERR_INVALID; }
    if (rows < 2 || n_mels < 1 || mel_sr < 1) { set_error("mask_mel_log_sr: rows=%d n_mels=%d mel_sr=%d", rows, n_mels, mel_sr); return HPSS_ERR_INVALID; }
    if (log_power && !(amin > 0.f)) { set_error("amin must be strictly positive (librosa.power_to_db)"); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    MelPlan* mp = nullptr;
    int rc = get_mel_plan(ctx, mel_sr, 2 * (rows - 1), n_mels, &mp);
    if (rc) return rc;
    return launch_mask_mel(ctx, batch, S, harm, perc, rows, mp->d_w, mp->d_band, mp->sweepable ? mp->d_sweep : nullptr,
                           mp->walkable ? mp->d_emit4 : nullptr, mp->walkable ? mp->d_sweep_w : nullptr, n_mels,
                           pre_square, log_power, amin, out, clip_max, (cudaStream_t)stream);
}

int hpss_topdb_clip(hpss_ctx* ctx, const hpss_batch* batch, float* out, int32_t rows_per_stream, int32_t n_streams,
                    const uint32_t* clip_max, float top_db, void* stream) {
    if (!ctx || !batch || !out || !clip_max) { set_error("topdb_clip: NULL argument"); return HPSS_ERR_INVALID; }
    if (top_db < 0.f) { set_error("top_db must be non-negative (librosa.power_to_db)"); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    return launch_topdb(ctx, batch, out, rows_per_stream, n_streams, clip_max, top_db, (cudaStream_t)stream);
}

int32_t hpss_feature_rows(const hpss_params* p) {
    if (!p) return -1;
    const int ns = feature_streams(p->feature);
    return ns * (feature_is_mel(p->feature) ? p->n_mels : p->n_fft / 2 + 1);
}

int hpss_featuregram(hpss_ctx* ctx, const hpss_batch* batch, const float* wave, const hpss_params* p, float* out,
                     void* stream) {
    if (!ctx || !batch || !wave || !out) { set_error("featuregram: NULL argument"); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    ScratchLease lease;
    int rc = lease.begin(ctx, (cudaStream_t)stream);
    if (rc) return rc;
    return featuregram_device(ctx, const_cast<hpss_batch*>(batch), wave, p, out, (cudaStream_t)stream);
}

int hpss_featuregram_from_spec(hpss_ctx* ctx, const hpss_batch* batch, const float* S, int32_t rows,
                               const hpss_params* p, float* out, void* stream) {
    if (!ctx || !batch || !S || !out) { set_error("featuregram_from_spec: NULL argument"); return HPSS_ERR_INVALID; }
    int rc = check_params(p);
    if (rc) return rc;
    if (rows < 2) { set_error("rows=%d", rows); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (p->feature == HPSS_FEAT_SPEC) {
        HPSS_CUDA(cudaMemcpyAsync(out, S, sizeof(float) * (size_t)rows * batch->frame_off[batch->n_clips],
                                  cudaMemcpyDeviceToDevice, st));
        return HPSS_OK;
    }
    if (feature_streams(p->feature) == 2) {    // librosa.util.softmask raises on negative input: flagged for hpss_ctx_check
        rc = launch_check(ctx, S, (int64_t)rows * batch->frame_off[batch->n_clips], 1, st);
        if (rc) return rc;
    }
    ScratchLease lease;
    rc = lease.begin(ctx, st);
    if (rc) return rc;
    Workspace w;
    rc = carve_workspace(ctx, batch, rows, false, feature_streams(p->feature) == 2, 0, &w, nullptr);
    if (rc) return rc;
    return features_from_spec(ctx, batch, S, rows, p, w.harm, w.perc, w.clip_max, out, st);
}

// ---- host-buffer pipeline ---------------------------------------------------------------------------------
// Clips are cut into chunks (never splitting a clip); chunk i+1 uploads while chunk i computes and chunk i-1
// downloads: three streams, two device slots.  Input is either the prepared float32 waveform or the decoded
// file (float32 or int16 PCM) with the signal preparation of N2 run on the device; output is the features
// in host memory and / or the raw moments of get_data_stats (the 8 KB a corpus pass really needs).
static int ensure_pipe_streams(hpss_ctx* ctx) {
    if (ctx->s_h2d) return HPSS_OK;
    HPSS_CUDA(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
    HPSS_CUDA(cudaStreamCreateWithFlags(&ctx->s_comp, cudaStreamNonBlocking));
    HPSS_CUDA(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        HPSS_CUDA(cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming));
        HPSS_CUDA(cudaEventCreateWithFlags(&ctx->ev_comp[i], cudaEventDisableTiming));
        HPSS_CUDA(cudaEventCreateWithFlags(&ctx->ev_d2h[i], cudaEventDisableTiming));
    }
    return HPSS_OK;
}

int hpss_pipeline_destroy(hpss_pipeline* pl) {
    if (!pl) return HPSS_OK;
    cudaSetDevice(pl->ctx->device);
    cudaDeviceSynchronize();
    for (auto* sb : pl->subs) if (sb) hpss_batch_destroy(sb);
    for (auto* pp : pl->prep_plans) prep_plan_free(pp);
    for (int i = 0; i < 2; ++i) if (pl->slot[i]) cudaFree(pl->slot[i]);
    if (pl->d_acc) cudaFree(pl->d_acc);
    if (pl->h_acc) cudaFreeHost(pl->h_acc);
    if (pl->d_class) cudaFree(pl->d_class);
    delete pl;
    return HPSS_OK;
}

int hpss_pipeline_create(hpss_ctx* ctx, const int64_t* clip_len, int32_t n_clips, const hpss_params* p,
                         int32_t pcm_format, int32_t prepare, int32_t fs, double alpha, double beta,
                         int32_t n_chunks_hint, hpss_pipeline** out) {
    if (!ctx || !out || n_clips < 0 || (n_clips > 0 && !clip_len)) { set_error("pipeline_create: bad arguments"); return HPSS_ERR_INVALID; }
    *out = nullptr;
    int rc = check_params(p);
    if (rc) return rc;
    if (pcm_format != HPSS_PCM_F32 && pcm_format != HPSS_PCM_S16) { set_error("pipeline_create: unknown pcm_format %d", pcm_format); return HPSS_ERR_INVALID; }
    if (!prepare && pcm_format != HPSS_PCM_F32) { set_error("pipeline_create: int16 PCM needs prepare != 0 (the STFT reads float32)"); return HPSS_ERR_INVALID; }
    if (prepare && fs < 1) { set_error("pipeline_create: fs=%d", fs); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    hpss_pipeline* pl = new hpss_pipeline();
    pl->ctx = ctx; pl->prm = *p; pl->n_clips = n_clips; pl->pcm_format = pcm_format; pl->prepare = prepare; pl->fs = fs;
    pl->alpha = alpha; pl->beta = beta;
    pl->rows_out = hpss_feature_rows(p);
    pl->in_len.assign(clip_len, clip_len + n_clips);
    pl->in_off.assign(n_clips + 1, 0);
    pl->wav_len.resize(n_clips);
    pl->frame_off.assign(n_clips + 1, 0);
    for (int c = 0; c < n_clips; ++c) {
        pl->in_off[c + 1] = pl->in_off[c] + clip_len[c];
        pl->wav_len[c] = prepare ? prep_out_length(clip_len[c], fs) : clip_len[c];
        if (pl->wav_len[c] < p->n_fft) {
            set_error("n_fft=%d is too large for input signal of length=%lld (clip %d)", p->n_fft, (long long)pl->wav_len[c], c);
            delete pl;
            return HPSS_ERR_SHORT_SIGNAL;
        }
        pl->frame_off[c + 1] = pl->frame_off[c] + 1 + (pl->wav_len[c] - p->n_fft) / p->hop_length;
    }
    // chunking: ~16 chunks, at least 1 M samples each, never splitting a clip;
    const int64_t total = pl->in_off[n_clips];
    const int n_target = n_chunks_hint > 0 ? n_chunks_hint : (knobs().host_chunks > 0 ? knobs().host_chunks : 16);
    // the last chunk is about half a chunk: its compute (and download) is the part of the pass no transfer hides
    const int64_t target = std::max<int64_t>(n_target >= 2 ? (int64_t)((double)total / (n_target - 0.5)) + 1 : total, 1 << 20);
    pl->cut.assign(1, 0);
    for (int c = 0; c < n_clips;) {
        int e = c;
        int64_t acc = 0;
        while (e < n_clips && (e == c || acc + clip_len[e] <= target)) { acc += clip_len[e]; ++e; }
        pl->cut.push_back(e);
        c = e;
    }
    const int n_chunks = (int)pl->cut.size() - 1;
    pl->subs.assign(n_chunks, nullptr);
    const int rows = p->n_fft / 2 + 1;
    const size_t in_elt = pcm_format == HPSS_PCM_S16 ? 2 : 4;
    for (int i = 0; i < n_chunks; ++i) {
        const int c0 = pl->cut[i], c1 = pl->cut[i + 1];
        rc = hpss_batch_from_samples(ctx, pl->wav_len.data() + c0, c1 - c0, p->n_fft, p->hop_length, &pl->subs[i]);
        if (rc) { hpss_pipeline_destroy(pl); return rc; }
        if (prepare) {
            PrepPlan* pp = nullptr;
            rc = prep_plan_build(ctx, clip_len + c0, c1 - c0, fs, p->win_length, p->hop_length, &pp);
            if (rc) { hpss_pipeline_destroy(pl); return rc; }
            pl->prep_plans.push_back(pp);
        }
        int64_t wav = 0;
        for (int c = c0; c < c1; ++c) wav += pl->wav_len[c];
        const int64_t frames = pl->frame_off[c1] - pl->frame_off[c0];
        pl->in_bytes = std::max(pl->in_bytes, align256((size_t)(pl->in_off[c1] - pl->in_off[c0]) * in_elt));
        pl->wav_bytes = std::max(pl->wav_bytes, align256((size_t)wav * 4));
        pl->out_bytes = std::max(pl->out_bytes, align256((size_t)pl->rows_out * (size_t)frames * 4));
        const size_t sz = align256((size_t)rows * (size_t)frames * sizeof(float));
        pl->ws_need = std::max(pl->ws_need, 3 * sz + align256(sizeof(uint32_t) * 2 * (size_t)std::max(1, c1 - c0)));
    }
    const size_t slot_bytes = (prepare ? pl->in_bytes : 0) + pl->wav_bytes + pl->out_bytes;
    for (int i = 0; i < 2 && n_chunks > 0; ++i) {
        cudaError_t e = cudaMalloc(&pl->slot[i], std::max<size_t>(slot_bytes, 256));
        if (e != cudaSuccess) { cudaGetLastError(); hpss_pipeline_destroy(pl); set_error("out of device memory: pipeline slot of %zu bytes", slot_bytes); return HPSS_ERR_NOMEM; }
    }
    *out = pl;
    return HPSS_OK;
}

int64_t hpss_pipeline_total_frames(const hpss_pipeline* pl) { return pl ? pl->frame_off[pl->n_clips] : 0; }
int32_t hpss_pipeline_n_chunks(const hpss_pipeline* pl) { return pl ? (int32_t)pl->subs.size() : 0; }
int hpss_pipeline_frame_offsets(const hpss_pipeline* pl, int64_t* o) {
    if (!pl || !o) { set_error("pipeline_frame_offsets: NULL argument"); return HPSS_ERR_INVALID; }
    memcpy(o, pl->frame_off.data(), sizeof(int64_t) * (pl->n_clips + 1));
    return HPSS_OK;
}

int hpss_pipeline_run(hpss_pipeline* pl, const void* pcm_host, float* feat_host, const int32_t* clip_class,
                      int32_t n_classes, double* moments_host) {
    if (!pl || !pcm_host) { set_error("pipeline_run: NULL argument"); return HPSS_ERR_INVALID; }
    if ((clip_class == nullptr) != (moments_host == nullptr)) { set_error("pipeline_run: clip_class_host and moments_host go together"); return HPSS_ERR_INVALID; }
    if (!feat_host && !moments_host) { set_error("pipeline_run: nothing to produce (feat_host and moments_host are NULL)"); return HPSS_ERR_INVALID; }
    hpss_ctx* ctx = pl->ctx;
    const hpss_params* p = &pl->prm;
    const int n_chunks = (int)pl->subs.size();
    if (n_chunks == 0) return HPSS_OK;
    HPSS_CUDA(cudaSetDevice(ctx->device));
    int rc = ensure_pipe_streams(ctx);
    if (rc) return rc;
    const int D = pl->rows_out;
    const bool want_mom = moments_host != nullptr;
    if (want_mom) {
        if (n_classes < 1 || n_classes > 8) { set_error("pipeline_run: n_classes=%d (1..8)", n_classes); return HPSS_ERR_INVALID; }
        for (int c = 0; c < pl->n_clips; ++c)
            if (clip_class[c] < 0 || clip_class[c] >= n_classes) { set_error("clip %d has class %d outside [0,%d)", c, clip_class[c], n_classes); return HPSS_ERR_INVALID; }
        const int n = n_classes * D + D + n_classes + 1;
        if (pl->acc_n < n) {
            if (pl->d_acc) { HPSS_CUDA(cudaFree(pl->d_acc)); pl->d_acc = nullptr; }
            if (pl->h_acc) { HPSS_CUDA(cudaFreeHost(pl->h_acc)); pl->h_acc = nullptr; }
            HPSS_CUDA(cudaMalloc(&pl->d_acc, sizeof(double) * n));
            HPSS_CUDA(cudaHostAlloc(&pl->h_acc, sizeof(double) * n, cudaHostAllocDefault));
            pl->acc_n = n;
        }
        if (!pl->d_class) HPSS_CUDA(cudaMalloc(&pl->d_class, sizeof(int32_t) * std::max(1, pl->n_clips)));
    }
    ScratchLease lease;                                   // the chunks share ctx->ws / prep_ws on s_comp
    rc = lease.begin(ctx, ctx->s_comp);
    if (rc) return rc;
    rc = ensure_workspace(ctx, pl->ws_need);
    if (rc) return rc;
    const int n = n_classes * D + D + n_classes + 1;
    if (want_mom) {
        HPSS_CUDA(cudaMemsetAsync(pl->d_acc, 0, sizeof(double) * n, ctx->s_comp));
        HPSS_CUDA(cudaMemcpyAsync(pl->d_class, clip_class, sizeof(int32_t) * pl->n_clips, cudaMemcpyHostToDevice, ctx->s_comp));
    }
    const size_t in_elt = pl->pcm_format == HPSS_PCM_S16 ? 2 : 4;
    const char* src = (const char*)pcm_host;
    for (int i = 0; i < n_chunks; ++i) {
        const int s = i & 1;
        const int c0 = pl->cut[i], c1 = pl->cut[i + 1];
        char* base = (char*)pl->slot[s];
        void* d_in = base;                                                    // raw PCM (prepare) or the waveform itself
        float* d_wave = pl->prepare ? (float*)(base + pl->in_bytes) : (float*)base;
        float* d_out = (float*)((char*)d_wave + pl->wav_bytes);
        const size_t in_b = (size_t)(pl->in_off[c1] - pl->in_off[c0]) * in_elt;
        const size_t os = (size_t)D * (size_t)(pl->frame_off[c1] - pl->frame_off[c0]);
        cudaError_t e = cudaSuccess;
        if (i >= 2) e = cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_comp[s], 0);      // input slot free
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(d_in, src + (size_t)pl->in_off[c0] * in_elt, in_b, cudaMemcpyHostToDevice, ctx->s_h2d);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_h2d[s], ctx->s_h2d);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->s_comp, ctx->ev_h2d[s], 0);
        if (e == cudaSuccess && i >= 2 && feat_host) e = cudaStreamWaitEvent(ctx->s_comp, ctx->ev_d2h[s], 0);   // out slot free
        if (e != cudaSuccess) { cudaDeviceSynchronize(); return cuda_fail(e, "host pipeline (upload)"); }
        if (pl->prepare) {
            rc = launch_prep_plan(ctx, pl->prep_plans[i], d_in, pl->pcm_format, pl->fs, p->win_length, p->hop_length,
                                  pl->alpha, pl->beta, d_wave, ctx->s_comp);
        } else {
            int64_t wav = 0;
            for (int c = c0; c < c1; ++c) wav += pl->wav_len[c];
            rc = launch_check(ctx, d_wave, wav, 0, ctx->s_comp);               // librosa.util.valid_audio
        }
        if (rc) { cudaDeviceSynchronize(); return rc; }
        MomentSink ms;
        if (want_mom) {
            ms.d_class = pl->d_class + c0; ms.n_classes = n_classes;
            ms.sum = pl->d_acc; ms.sumsq = pl->d_acc + (size_t)n_classes * D;
            ms.count = ms.sumsq + D; ms.nonfinite = ms.count + n_classes;
        }
        rc = featuregram_device(ctx, pl->subs[i], d_wave, p, d_out, ctx->s_comp, want_mom ? &ms : nullptr);
        if (rc) { cudaDeviceSynchronize(); return rc; }
        e = cudaEventRecord(ctx->ev_comp[s], ctx->s_comp);
        if (feat_host) {
            if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_comp[s], 0);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(feat_host + (size_t)D * (size_t)pl->frame_off[c0], d_out, os * sizeof(float),
                                    cudaMemcpyDeviceToHost, ctx->s_d2h);
            if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_d2h[s], ctx->s_d2h);
        }
        if (e != cudaSuccess) { cudaDeviceSynchronize(); return cuda_fail(e, "host pipeline (download)"); }
    }
    cudaError_t e = cudaSuccess;
    if (want_mom) e = cudaMemcpyAsync(pl->h_acc, pl->d_acc, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->s_comp);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->s_comp);
    if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_flags, 0, sizeof(uint32_t), ctx->s_comp);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->s_comp);
    if (e == cudaSuccess && feat_host) e = cudaStreamSynchronize(ctx->s_d2h);
    if (e != cudaSuccess) return cuda_fail(e, "host pipeline (sync)");
    if (want_mom)
        for (int i = 0; i < n; ++i) moments_host[i] += pl->h_acc[i];
    if (ctx->h_flags[0] & 1u) { set_error("Audio buffer is not finite everywhere"); return HPSS_ERR_NONFINITE; }
    return HPSS_OK;
}

// Prepared float32 waveform in host memory -> features in host memory (the pipeline is cached in the batch).
int hpss_featuregram_host(hpss_ctx* ctx, const hpss_batch* batch, const float* wave_host, const hpss_params* p,
                          float* out_host) {
    if (!ctx || !batch || !wave_host || !out_host) { set_error("featuregram_host: NULL argument"); return HPSS_ERR_INVALID; }
    int rc = check_params(p);
    if (rc) return rc;
    if (!batch->has_samples) { set_error("featuregram_host needs a batch built from sample lengths"); return HPSS_ERR_INVALID; }
    if (batch->n_fft != p->n_fft || batch->hop != p->hop_length) {
        set_error("batch was laid out for n_fft=%d hop=%d, params say %d / %d", batch->n_fft, batch->hop, p->n_fft, p->hop_length);
        return HPSS_ERR_INVALID;
    }
    hpss_batch* pb = const_cast<hpss_batch*>(batch);
    hpss_pipeline* pl = nullptr;
    {
        std::lock_guard<std::mutex> lk(pb->mu);
        if (pb->host_pipe && memcmp(&pb->host_pipe->prm, p, sizeof(*p)) != 0) { hpss_pipeline_destroy(pb->host_pipe); pb->host_pipe = nullptr; }
        if (!pb->host_pipe) {
            std::vector<int64_t> len(batch->n_clips);
            for (int c = 0; c < batch->n_clips; ++c) len[c] = batch->sample_off[c + 1] - batch->sample_off[c];
            rc = hpss_pipeline_create(ctx, len.data(), batch->n_clips, p, HPSS_PCM_F32, 0, 16000, 0.0, 0.0, 0, &pb->host_pipe);
            if (rc) return rc;
        }
        pl = pb->host_pipe;
    }
    return hpss_pipeline_run(pl, wave_host, out_host, nullptr, 0, nullptr);
}

// class per clip on the device, kept across calls (re-uploaded only when the classes change)
static int upload_classes(hpss_batch* b, const int32_t* clip_class, int n_given, int n_classes, cudaStream_t st) {
    if (n_given != b->n_clips) { set_error("clip_class has %d entries, the batch has %d clips", n_given, b->n_clips); return HPSS_ERR_INVALID; }
    for (int c = 0; c < b->n_clips; ++c)
        if (clip_class[c] < 0 || clip_class[c] >= n_classes) { set_error("clip %d has class %d outside [0,%d)", c, clip_class[c], n_classes); return HPSS_ERR_INVALID; }
    std::lock_guard<std::mutex> lk(b->mu);
    if (!b->d_clip_class && b->n_clips > 0) HPSS_CUDA(cudaMalloc(&b->d_clip_class, sizeof(int32_t) * b->n_clips));
    if (b->n_clips > 0 && (b->h_clip_class.size() != (size_t)b->n_clips ||
                           memcmp(b->h_clip_class.data(), clip_class, sizeof(int32_t) * b->n_clips) != 0)) {
        b->h_clip_class.assign(clip_class, clip_class + b->n_clips);
        HPSS_CUDA(cudaMemcpyAsync(b->d_clip_class, b->h_clip_class.data(), sizeof(int32_t) * b->n_clips, cudaMemcpyHostToDevice, st));
    }
    return HPSS_OK;
}

int hpss_featuregram_moments(hpss_ctx* ctx, const hpss_batch* batch, const float* wave, const hpss_params* p,
                             float* out, const int32_t* clip_class, int32_t n_clips, int32_t n_classes, double* sum,
                             double* sumsq, double* count, double* nonfinite, void* stream) {
    if (!ctx || !batch || !wave || !out || !clip_class || !sum || !sumsq || !count || !nonfinite) { set_error("featuregram_moments: NULL argument"); return HPSS_ERR_INVALID; }
    if (n_classes < 1) { set_error("featuregram_moments: n_classes=%d", n_classes); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    hpss_batch* b = const_cast<hpss_batch*>(batch);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = upload_classes(b, clip_class, n_clips, n_classes, st);
    if (rc) return rc;
    MomentSink ms;
    ms.d_class = b->d_clip_class; ms.n_classes = n_classes;
    ms.sum = sum; ms.sumsq = sumsq; ms.count = count; ms.nonfinite = nonfinite;
    ScratchLease lease;
    rc = lease.begin(ctx, st);
    if (rc) return rc;
    return featuregram_device(ctx, b, wave, p, out, st, &ms);
}

int hpss_topdb_moments(hpss_ctx* ctx, const hpss_batch* batch, float* out, int32_t rows_per_stream, int32_t n_streams,
                       const uint32_t* clip_max, float top_db, const int32_t* clip_class, int32_t n_clips,
                       int32_t n_classes, double* sum, double* sumsq, double* count, double* nonfinite, void* stream) {
    if (!ctx || !batch || !out || !clip_max || !clip_class || !sum || !sumsq || !count || !nonfinite) { set_error("topdb_moments: NULL argument"); return HPSS_ERR_INVALID; }
    if (top_db < 0.f || rows_per_stream < 1 || n_streams < 1 || n_classes < 1) { set_error("topdb_moments: bad argument"); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    hpss_batch* b = const_cast<hpss_batch*>(batch);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = upload_classes(b, clip_class, n_clips, n_classes, st);
    if (rc) return rc;
    return launch_topdb_moments(ctx, batch, out, rows_per_stream, n_streams, clip_max, top_db, b->d_clip_class, n_classes,
                                sum, sumsq, count, nonfinite, st);
}

int hpss_moments(hpss_ctx* ctx, const hpss_batch* batch, const float* feat, int32_t D, const int32_t* clip_class,
                 int32_t n_clips, int32_t n_classes, double* sum, double* sumsq, double* count, double* nonfinite,
                 void* stream) {
    if (!ctx || !batch || !feat || !clip_class || !sum || !sumsq || !count || !nonfinite) { set_error("moments: NULL argument"); return HPSS_ERR_INVALID; }
    if (D < 1 || n_classes < 1) { set_error("moments: D=%d n_classes=%d", D, n_classes); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    hpss_batch* b = const_cast<hpss_batch*>(batch);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = upload_classes(b, clip_class, n_clips, n_classes, st);
    if (rc) return rc;
    return launch_moments(ctx, batch, feat, D, b->d_clip_class, n_classes, sum, sumsq, count, nonfinite, st);
}

// ---- N2: signal preparation -----------------------------------------------------------------------------------
int64_t hpss_prep_out_length(int64_t n_samples, int32_t fs) { return fs > 0 ? prep_out_length(n_samples, fs) : 0; }
int64_t hpss_prep_num_frames(int64_t n_samples, int32_t win, int32_t hop) { return prep_num_frames(n_samples, win, hop); }

int hpss_prep_signals(hpss_ctx* ctx, const void* pcm, int32_t pcm_format, const int64_t* clip_len, int32_t n_clips,
                      int32_t fs, int32_t win, int32_t hop, double alpha, double beta, float* out,
                      int32_t* frame_marker, uint8_t* sample_marker, int32_t* n_sil, void* stream) {
    if (!ctx || !pcm || !out || n_clips < 0 || (n_clips > 0 && !clip_len)) { set_error("prep_signals: NULL argument"); return HPSS_ERR_INVALID; }
    if (pcm_format != HPSS_PCM_F32 && pcm_format != HPSS_PCM_S16) { set_error("prep_signals: unknown pcm_format %d", pcm_format); return HPSS_ERR_INVALID; }
    if (fs < 1 || win < 1 || hop < 1) { set_error("prep_signals: fs=%d win_length=%d hop_length=%d", fs, win, hop); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    ScratchLease lease;
    int rc = lease.begin(ctx, (cudaStream_t)stream);
    if (rc) return rc;
    return launch_prep(ctx, pcm, pcm_format, clip_len, n_clips, fs, win, hop, alpha, beta, out, frame_marker,
                       sample_marker, n_sil, (cudaStream_t)stream);
}

int hpss_mix_signals(hpss_ctx* ctx, const float* sp, const int64_t* sp_len, const float* mu, const int64_t* mu_len,
                     const double* target_db, int32_t n_pairs, float* out, void* stream) {
    if (!ctx || !sp || !mu || !out || n_pairs < 0 || (n_pairs > 0 && (!sp_len || !mu_len || !target_db))) { set_error("mix_signals: NULL argument"); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    ScratchLease lease;
    int rc = lease.begin(ctx, (cudaStream_t)stream);
    if (rc) return rc;
    return launch_mix(ctx, sp, sp_len, mu, mu_len, target_db, n_pairs, out, (cudaStream_t)stream);
}

int hpss_validate_audio(hpss_ctx* ctx, const float* wave, int64_t n, void* stream) {
    if (!ctx || (!wave && n > 0)) { set_error("validate_audio: NULL argument"); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    return launch_check(ctx, wave, n, 0, (cudaStream_t)stream);
}

int hpss_validate_nonneg(hpss_ctx* ctx, const float* x, int64_t n, void* stream) {
    if (!ctx || (!x && n > 0)) { set_error("validate_nonneg: NULL argument"); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    return launch_check(ctx, x, n, 1, (cudaStream_t)stream);
}

// Synchronises `stream` and reports what the kernels flagged since the last check.
int hpss_ctx_check(hpss_ctx* ctx, void* stream) {
    if (!ctx) { set_error("ctx_check: NULL context"); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t flags = 0;
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        HPSS_CUDA(cudaMemcpyAsync(ctx->h_flags + 1, ctx->d_flags, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        HPSS_CUDA(cudaMemsetAsync(ctx->d_flags, 0, sizeof(uint32_t), st));
        HPSS_CUDA(cudaStreamSynchronize(st));
        flags = ctx->h_flags[1];
    }
    if (flags & 1u) { set_error("Audio buffer is not finite everywhere"); return HPSS_ERR_NONFINITE; }
    if (flags & 2u) { set_error("X and X_ref must be non-negative"); return HPSS_ERR_NEGATIVE; }
    return HPSS_OK;
}

int hpss_stats_finalize(const double* sum, const double* sumsq, const double* count, int32_t D, int32_t n_classes,
                        float* mean_out, float* stdev_out) {
    if (!sum || !sumsq || !count || !mean_out || !stdev_out || D < 1 || n_classes < 1) { set_error("stats_finalize: bad argument"); return HPSS_ERR_INVALID; }
    double N = 0;
    for (int k = 0; k < n_classes; ++k) N += count[k];
    for (int d = 0; d < D; ++d) {
        double mu = 0, tot = 0;
        for (int k = 0; k < n_classes; ++k) {
            mu += sum[(size_t)k * D + d] / (count[k] + 1e-10);   // class mean (:530-532)
            tot += sum[(size_t)k * D + d];
        }
        mu /= (double)n_classes;                                  // unweighted mean of class means (:533-536)
        const double ss = sumsq[d] - 2.0 * mu * tot + N * mu * mu;   // sum (x - mu)^2
        mean_out[d] = (float)mu;
        stdev_out[d] = (float)sqrt(std::max(ss, 0.0) / (N - 1.0));
    }
    return HPSS_OK;
}

int hpss_scale_data(hpss_ctx* ctx, const hpss_batch* batch, const float* feat, int32_t D, const float* mean,
                    const float* stdev, double eps, double* out, void* stream) {
    if (!ctx || !batch || !feat || !mean || !stdev || !out) { set_error("scale_data: NULL argument"); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    return launch_scale(ctx, batch, feat, D, mean, stdev, eps, out, nullptr, (cudaStream_t)stream);
}

int hpss_scale_data_f32(hpss_ctx* ctx, const hpss_batch* batch, const float* feat, int32_t D, const float* mean,
                        const float* stdev, float* out, void* stream) {
    if (!ctx || !batch || !feat || !mean || !stdev || !out) { set_error("scale_data_f32: NULL argument"); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    return launch_scale(ctx, batch, feat, D, mean, stdev, 0.0, nullptr, out, (cudaStream_t)stream);
}

int hpss_dct_mfcc(hpss_ctx* ctx, const hpss_batch* batch, const float* feat, int32_t rows_per_stream, int32_t n_streams,
                  int32_t n_mfcc, float* out, void* stream) {
    if (!ctx || !batch || !feat || !out || feat == out) { set_error("dct_mfcc: NULL or aliased argument"); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    return launch_dct(ctx, batch, feat, rows_per_stream, n_streams, n_mfcc, out, (cudaStream_t)stream);
}

int hpss_dct_basis(int32_t n_mels, int32_t n_mfcc, float* out) {
    if (!out || n_mels < 1 || n_mfcc < 1 || n_mfcc > n_mels) { set_error("dct_basis: need 1 <= n_mfcc <= n_mels"); return HPSS_ERR_INVALID; }
    std::vector<float> t((size_t)n_mels * n_mfcc);
    build_dct_basis_t(n_mels, n_mfcc, n_mfcc, t.data());
    for (int k = 0; k < n_mfcc; ++k)
        for (int m = 0; m < n_mels; ++m) out[(size_t)k * n_mels + m] = t[(size_t)m * n_mfcc + k];
    return HPSS_OK;
}

int hpss_row_standardize(hpss_ctx* ctx, const hpss_batch* batch, float* feat, int32_t D, void* stream) {
    if (!ctx || !batch || !feat || D < 1) { set_error("row_standardize: bad argument"); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    return launch_row_standardize(ctx, batch, feat, D, (cudaStream_t)stream);
}

int64_t hpss_num_patches(int64_t n_frames, int32_t W, int32_t shift) {
    if (W < 1 || shift < 1) return 0;
    const int64_t half = W / 2;
    const int64_t a = half, b = n_frames - half;
    return b > a ? (b - a + shift - 1) / shift : 0;
}

// patches of one clip of T frames, including the tiling of clips shorter than the patch (lib/preprocessing.py:139-142)
int64_t hpss_num_patches_tiled(int64_t T, int32_t W, int32_t shift) {
    if (T < 1 || W < 1 || shift < 1) return 0;
    int64_t Tt = T;
    if (T < W) Tt = T * (W / T + 1);          // FV is appended to itself while its length is <= W
    return hpss_num_patches(Tt, W, shift);
}

static int patch_offsets(hpss_batch* b, int W, int shift, const std::vector<int64_t>** host, const int64_t** dev) {
    std::lock_guard<std::mutex> lk(b->mu);
    auto key = std::make_pair(W, shift);
    auto it = b->patch_offs.find(key);
    if (it == b->patch_offs.end()) {
        std::vector<int64_t> off(b->n_clips + 1, 0);
        for (int c = 0; c < b->n_clips; ++c)
            off[c + 1] = off[c] + hpss_num_patches_tiled(b->frame_off[c + 1] - b->frame_off[c], W, shift);
        int64_t* d = nullptr;
        HPSS_CUDA(cudaMalloc(&d, sizeof(int64_t) * off.size()));
        HPSS_CUDA(cudaMemcpy(d, off.data(), sizeof(int64_t) * off.size(), cudaMemcpyHostToDevice));
        it = b->patch_offs.emplace(key, std::make_pair(std::move(off), d)).first;
    }
    *host = &it->second.first;
    *dev = it->second.second;
    return HPSS_OK;
}

int hpss_patch_offsets(const hpss_batch* batch, int32_t W, int32_t shift, int64_t* out) {
    if (!batch || !out || W < 1 || shift < 1) { set_error("patch_offsets: bad argument"); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(batch->ctx->device));
    const std::vector<int64_t>* h; const int64_t* d;
    int rc = patch_offsets(const_cast<hpss_batch*>(batch), W, shift, &h, &d);
    if (rc) return rc;
    memcpy(out, h->data(), sizeof(int64_t) * h->size());
    return HPSS_OK;
}

int hpss_patch_tensor(hpss_ctx* ctx, const hpss_batch* batch, float* feat, int32_t D, int32_t standardize, int32_t row0,
                      int32_t n_rows, int32_t W, int32_t shift, int32_t time_major, int32_t out_f64, void* out,
                      void* stream) {
    if (!ctx || !batch || !feat || !out) { set_error("patch_tensor: NULL argument"); return HPSS_ERR_INVALID; }
    if (D < 1 || row0 < 0 || n_rows < 1 || row0 + n_rows > D || W < 1 || shift < 1) { set_error("patch_tensor: bad shape arguments"); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const std::vector<int64_t>* h; const int64_t* d;
    int rc = patch_offsets(const_cast<hpss_batch*>(batch), W, shift, &h, &d);
    if (rc) return rc;
    if (standardize) {
        rc = launch_row_standardize(ctx, batch, feat, D, st);
        if (rc) return rc;
    }
    return launch_patch_tensor(batch, feat, 0, d, h->back(), D, row0, n_rows, W, shift, time_major, out_f64, out, st);
}

int hpss_patch_tensor_f64(hpss_ctx* ctx, const hpss_batch* batch, const double* feat, int32_t D, int32_t row0,
                          int32_t n_rows, int32_t W, int32_t shift, int32_t time_major, double* out, void* stream) {
    if (!ctx || !batch || !feat || !out) { set_error("patch_tensor_f64: NULL argument"); return HPSS_ERR_INVALID; }
    if (D < 1 || row0 < 0 || n_rows < 1 || row0 + n_rows > D || W < 1 || shift < 1) { set_error("patch_tensor_f64: bad shape arguments"); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    const std::vector<int64_t>* h; const int64_t* d;
    int rc = patch_offsets(const_cast<hpss_batch*>(batch), W, shift, &h, &d);
    if (rc) return rc;
    return launch_patch_tensor(batch, feat, 1, d, h->back(), D, row0, n_rows, W, shift, time_major, 1, out, (cudaStream_t)stream);
}

int hpss_row_nonfinite(hpss_ctx* ctx, const hpss_batch* batch, const float* feat, int32_t D, uint8_t* flags, void* stream) {
    if (!ctx || !batch || !feat || !flags || D < 1) { set_error("row_nonfinite: bad argument"); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    return launch_row_nonfinite(ctx, batch, feat, D, flags, (cudaStream_t)stream);
}

int hpss_patch_statistics(hpss_ctx* ctx, const double* patches, int64_t n_patches, int32_t n_feat, int32_t n_frames,
                          int32_t stat, int32_t axis, double* out, void* stream) {
    if (!ctx || !patches || !out || n_patches < 0 || n_feat < 1 || n_frames < 1) { set_error("patch_statistics: bad argument"); return HPSS_ERR_INVALID; }
    if (stat < 0 || stat > 3 || (axis != 0 && axis != 1)) { set_error("patch_statistics: stat in 0..3 (mean, variance, skew, kurtosis), axis 0 or 1"); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    return launch_patch_stats(ctx, patches, n_patches, n_feat, n_frames, stat, axis == 0 ? 1 : 0, out, (cudaStream_t)stream);
}

int hpss_extract_patches(hpss_ctx* ctx, const float* feat, int32_t D, int64_t T, int32_t W, int32_t shift,
                         double* out, void* stream) {
    if (!ctx || !feat || !out || D < 1 || W < 1 || shift < 1) { set_error("extract_patches: bad argument"); return HPSS_ERR_INVALID; }
    if (T < W) { set_error("extract_patches: %lld frames < patch_size %d (caller tiles the clip first, lib/preprocessing.py:139-142)", (long long)T, W); return HPSS_ERR_INVALID; }
    HPSS_CUDA(cudaSetDevice(ctx->device));
    return launch_patches(feat, D, T, W, shift, hpss_num_patches(T, W, shift), out, (cudaStream_t)stream);
}

}  // extern "C"
