// svad_segments.cpp -- speech-segment extraction from per-chunk probabilities (host C++), the
// post-processing half of get_speech_timestamps for one stream or a whole batch of streams.
//
// Reference behaviour reproduced (not its code): src/silero_vad/utils_vad.py:315-319 (thresholds in
// samples are real-valued), :338-426 (hysteresis automaton: enter at p >= threshold, tentative end at
// p < neg_threshold, close after min_silence, drop short segments, split over-long speech at the longest
// recorded silence or at the last long-enough one), :428-440 (padding / splitting the gap between
// neighbours).  Python-float semantics are kept by doing all threshold arithmetic in double and by
// comparing the float32 probability after widening, exactly what `.item()` hands the reference.
// The native twin in the reference tree is examples/cpp/silero-vad-onnx.cpp:199-331.
#include <cmath>
#include <cstdint>
#include <functional>
#include <thread>
#include <vector>

#include "../../include/silero_vad_b200.h"

namespace {

struct Seg { int64_t start, end; };

inline int64_t floordiv2(int64_t v) { return (v >= 0) ? v / 2 : -((-v + 1) / 2); }

int64_t segments_one(const float* probs, int64_t T, int64_t audio_len, const svad_segment_params& p, std::vector<Seg>& out) {
    out.clear();
    const int64_t w = p.sampling_rate == 16000 ? 512 : 256;
    const double sr = (double)p.sampling_rate;
    const double min_speech = sr * p.min_speech_duration_ms / 1000.0;
    const double pad = sr * p.speech_pad_ms / 1000.0;
    const double max_speech = sr * p.max_speech_duration_s - (double)w - 2.0 * pad;
    const double min_sil = sr * p.min_silence_duration_ms / 1000.0;
    const double min_sil_max = sr * p.min_silence_at_max_speech_ms / 1000.0;
    const double thr = p.threshold;
    const double neg = std::isnan(p.neg_threshold) ? std::fmax(thr - 0.15, 0.01) : p.neg_threshold;

    bool triggered = false, have_cur = false;
    int64_t cur_start = 0, temp_end = 0, prev_end = 0, next_start = 0;
    struct Cand { int64_t end, dur; };
    std::vector<Cand> cands;  // silences inside the running segment that are long enough to cut at

    auto reset_marks = [&] { prev_end = next_start = temp_end = 0; cands.clear(); };

    for (int64_t i = 0; i < T; i++) {
        const double pr = (double)probs[i];
        const int64_t s = w * i;
        if (pr >= thr && temp_end) {
            const int64_t sil = s - temp_end;
            if ((double)sil > min_sil_max) cands.push_back({temp_end, sil});
            temp_end = 0;
            if (next_start < prev_end) next_start = s;
        }
        if (pr >= thr && !triggered) {
            triggered = true; have_cur = true; cur_start = s;
            continue;
        }
        if (triggered && (double)(s - cur_start) > max_speech) {
            if (p.use_max_poss_sil_at_max_speech && !cands.empty()) {
                size_t best = 0;
                for (size_t k = 1; k < cands.size(); k++)
                    if (cands[k].dur > cands[best].dur) best = k;  // first maximum, like max(..., key=)
                prev_end = cands[best].end;
                const int64_t dur = cands[best].dur;
                out.push_back({cur_start, prev_end});
                have_cur = false;
                next_start = prev_end + dur;
                if (next_start < prev_end + s) { have_cur = true; cur_start = next_start; }
                else triggered = false;
                reset_marks();
            } else if (prev_end) {
                out.push_back({cur_start, prev_end});
                have_cur = false;
                if (next_start < prev_end) triggered = false;
                else { have_cur = true; cur_start = next_start; }
                reset_marks();
            } else {
                out.push_back({cur_start, s});
                have_cur = false;
                reset_marks();
                triggered = false;
                continue;
            }
        }
        if (pr < neg && triggered) {
            if (!temp_end) temp_end = s;
            const int64_t sil_now = s - temp_end;
            if (!p.use_max_poss_sil_at_max_speech && (double)sil_now > min_sil_max) prev_end = temp_end;
            if ((double)sil_now < min_sil) continue;
            if ((double)(temp_end - cur_start) > min_speech) out.push_back({cur_start, temp_end});
            have_cur = false;
            reset_marks();
            triggered = false;
            continue;
        }
    }
    if (have_cur && (double)(audio_len - cur_start) > min_speech) out.push_back({cur_start, audio_len});

    const size_t ns = out.size();
    for (size_t i = 0; i < ns; i++) {
        if (i == 0) out[0].start = (int64_t)std::fmax(0.0, (double)out[0].start - pad);
        if (i + 1 != ns) {
            const int64_t gap = out[i + 1].start - out[i].end;
            if ((double)gap < 2.0 * pad) {
                out[i].end += floordiv2(gap);
                const int64_t ns_ = out[i + 1].start - floordiv2(gap);
                out[i + 1].start = ns_ > 0 ? ns_ : 0;
            } else {
                out[i].end = (int64_t)std::fmin((double)audio_len, (double)out[i].end + pad);
                out[i + 1].start = (int64_t)std::fmax(0.0, (double)out[i + 1].start - pad);
            }
        } else {
            out[i].end = (int64_t)std::fmin((double)audio_len, (double)out[i].end + pad);
        }
    }
    return (int64_t)ns;
}

}  // namespace

extern "C" void svad_segment_params_default(svad_segment_params* p) {
    if (!p) return;
    p->sampling_rate = 16000;
    p->threshold = 0.5;
    p->neg_threshold = NAN;
    p->min_speech_duration_ms = 250.0;
    p->max_speech_duration_s = INFINITY;
    p->min_silence_duration_ms = 100.0;
    p->speech_pad_ms = 30.0;
    p->min_silence_at_max_speech_ms = 98.0;
    p->use_max_poss_sil_at_max_speech = 1;
}

// Rows are independent streams: they are cut into contiguous ranges, one host thread per range (each thread keeps its
// segments in a private vector), then the ranges are stitched in row order.  Threads: hardware concurrency, capped at 32
// and at one per 64 rows (small batches stay single-threaded: thread start-up costs more than 64 automata).
extern "C" int svad_speech_segments(const float* probs, int64_t B, int64_t T, int64_t ldp, const int64_t* audio_len,
                                    const svad_segment_params* p, int64_t* seg_offsets, int64_t* seg_bounds, int64_t cap,
                                    int64_t* n_total) {
    if (!p || (B > 0 && (!probs || !audio_len || !seg_offsets)) || B < 0 || T < 0 || ldp < T || !n_total) return SVAD_EINVAL;
    if (p->sampling_rate != 16000 && p->sampling_rate != 8000) return SVAD_EINVAL;
    const int64_t w = p->sampling_rate == 16000 ? 512 : 256;
    int nthr = (int)std::thread::hardware_concurrency();
    if (nthr < 1) nthr = 1;
    if (nthr > 32) nthr = 32;
    if ((int64_t)nthr > (B + 63) / 64) nthr = (int)((B + 63) / 64);
    if (nthr < 1) nthr = 1;
    struct Range { int64_t b0, b1; std::vector<Seg> segs; std::vector<int64_t> count; };
    std::vector<Range> ranges((size_t)nthr);
    auto work = [&](Range& r) {
        std::vector<Seg> one;
        r.count.assign((size_t)(r.b1 - r.b0), 0);
        for (int64_t b = r.b0; b < r.b1; b++) {
            // a stream shorter than T chunks only has ceil(len / w) meaningful probabilities
            int64_t Tb = (audio_len[b] + w - 1) / w;
            if (Tb > T) Tb = T;
            r.count[(size_t)(b - r.b0)] = segments_one(probs + b * ldp, Tb, audio_len[b], *p, one);
            r.segs.insert(r.segs.end(), one.begin(), one.end());
        }
    };
    try {
        for (int i = 0; i < nthr; i++) { ranges[(size_t)i].b0 = B * i / nthr; ranges[(size_t)i].b1 = B * (i + 1) / nthr; }
        std::vector<std::thread> pool;
        for (int i = 1; i < nthr; i++) pool.emplace_back(work, std::ref(ranges[(size_t)i]));
        work(ranges[0]);
        for (auto& t : pool) t.join();
    } catch (...) {
        return SVAD_ENOMEM;
    }
    int64_t n = 0;
    for (const Range& r : ranges) {
        size_t k = 0;
        for (int64_t b = r.b0; b < r.b1; b++) {
            seg_offsets[b] = n;
            for (int64_t c = 0; c < r.count[(size_t)(b - r.b0)]; c++, k++, n++)
                if (seg_bounds && n < cap) { seg_bounds[2 * n] = r.segs[k].start; seg_bounds[2 * n + 1] = r.segs[k].end; }
        }
    }
    if (B > 0) seg_offsets[B] = n;
    *n_total = n;
    return SVAD_OK;
}
