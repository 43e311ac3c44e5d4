// host_purity.cpp — tumor purity from the per-position products of the two extract passes: the host math that consumes
// lps_extract_normal / lps_extract_tumor.  Restates TumorPurityEstimator::estimateTumorPurity
// (reference src/somatic_haplotag/TumorPurityEstimator.cpp:31-84) with its stages:
//   LCVF filters            buildPurityFeatureValueVec   :92-149   (thresholds are `float` constants compared with doubles, .h:280-284)
//   histogram + smoothing   Histogram                    :426-630  (cumulative percentage, Gaussian sigma 0.5 -> 5 taps, edge clamping)
//   peaks / valley          PeakProcessor                :632-1064 (peaks >= 5 % of the maximum, min distance 2, main / secondary peak,
//                                                                   lowest valley, 0.3 cumulative-percentage and 0.7 height limits)
//   box plot                statisticPurityData          :281-343  (linear-interpolated quartiles, 1.5 IQR whiskers, one outlier round)
//   model                   :65  purity = -3.3454 m + 14.7747 q + 4.0344 m^2 - 13.7777 m q - 5.2434 q^2 + 0.3058
// Pure host code: no device work, no context.  All arithmetic in the reference's own types so that the result is bit-identical.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <stdexcept>
#include <vector>
#include "../../include/lps.h"

namespace {

struct Feature { int idx; double ratio; int count; };
struct Bin { double count; double percentage; };
enum Trend { NONE_PT = 0, UP, DOWN, FLAG };
struct Peak { size_t index; double height; Trend left, right; bool main; };
struct Valley { size_t index; double height; double percentage; };

struct Histo {
    std::vector<Bin> h;
    size_t total = 0;
    double max_height = 0;
    size_t first = 0, last = 0;

    void statistics() {                                                      // Histogram::calculateStatistics :478-522
        double acc = 0.0;
        bool seen = false;
        for (size_t i = 0; i < h.size(); i++) {
            acc += (double)h[i].count / (double)total;
            h[i].percentage = acc;
            if (h[i].count > max_height) max_height = h[i].count;
            if (!seen && h[i].count > 0) { first = i; seen = true; }
            if (h[i].count > 0) last = i;
        }
        if (max_height == 0) throw std::runtime_error("max_height is 0 in histogram");
        h.resize(last + 1);
    }
};

std::vector<double> gaussian_kernel(double sigma) {                            // Histogram::createGaussianKernel :578-616
    int size = (int)(6 * sigma + 1);
    if (size % 2 == 0) size += 1;
    std::vector<double> k((size_t)size);
    const int half = size / 2;
    double sum = 0.0;
    for (int i = 0; i < size; i++) {
        const double x = i - half;
        k[(size_t)i] = std::exp(-0.5 * (x / sigma) * (x / sigma));
        sum += k[(size_t)i];
    }
    for (double &v : k) v /= sum;
    return k;
}

Peak get_peak(const std::vector<Peak> &peaks, size_t index, int offset) {      // PeakProcessor::getPeak :1041-1060
    for (size_t i = 0; i < peaks.size(); i++)
        if (peaks[i].index == index) return peaks.at(i + offset);
    throw std::runtime_error("Peak not found");
}

bool lowest_valley(const std::vector<Bin> &h, size_t start, size_t end, Valley &v) {   // PeakProcessor::findLowestValley :911-942
    if (start >= end || end > h.size()) return false;
    bool found = false;
    v.height = INT_MAX;
    for (size_t i = start + 1; i < end - 1; i++)
        if (h[i].count < h[i - 1].count && h[i].count < h[i + 1].count && (!found || h[i].count < v.height)) {
            v.index = i; v.height = h[i].count; v.percentage = h[i].percentage; found = true;
        }
    return found;
}

int valley_threshold(const std::vector<Feature> &feat) {                       // findBimodalValleyThreshold :158-233
    Histo hist;
    hist.h.assign(1000, Bin{0, 0.0});
    hist.total = feat.size();
    for (const Feature &f : feat) {                                           // Histogram::buildHistogram :443-476
        const size_t c = (size_t)f.count;
        if (c >= hist.h.size()) {
            const size_t grown = hist.h.size() * 2;
            if (grown >= 1000000) throw std::overflow_error("read count exceeds maximum histogram size");
            hist.h.resize(grown, Bin{0, 0.0});
        }
        hist.h[c].count++;
    }
    hist.statistics();
    Histo sm = hist;                                                          // getSmoothedHistogram(0.5) -> applyGaussianFilter :524-576
    {
        const std::vector<double> k = gaussian_kernel(0.5);
        const std::vector<Bin> tmp = sm.h;
        const size_t half = k.size() / 2;
        for (size_t i = 0; i < sm.h.size(); i++) {
            double s = 0.0;
            for (size_t j = 0; j < k.size(); j++) {
                size_t idx = 0;
                if (i + j >= half) { idx = i + j - half; if (idx >= sm.h.size()) idx = sm.h.size() - 1; }
                s += tmp[idx].count * k[j];
            }
            if (!std::isfinite(s)) throw std::runtime_error("invalid smoothed value");
            sm.h[i].count = s;
        }
        sm.statistics();
    }
    const std::vector<Bin> &h = sm.h;
    const double max_height = sm.max_height;
    const double peak_thr = (double)std::max((size_t)((double)max_height * 0.05), (size_t)1);
    std::vector<Peak> peaks;                                                  // findPeaks :649-696
    for (size_t i = 0; i < h.size(); i++) {
        bool is_peak = false;
        if (h[i].count < peak_thr) continue;
        else if (i == 0 && i != h.size() - 1) is_peak = h[i].count > h[i + 1].count;
        else if (i == h.size() - 1 && i != 0) is_peak = h[i].count > h[i - 1].count;
        else if (h.size() > 1) is_peak = h[i].count > h[i - 1].count && h[i].count > h[i + 1].count;
        if (is_peak) peaks.push_back(Peak{i, h[i].count, NONE_PT, NONE_PT, false});
    }
    if (peaks.empty()) throw std::runtime_error("No peaks found");
    for (size_t i = 0; peaks.size() >= 2 && i < peaks.size() - 1;) {          // removeClosePeaks(2) :698-726
        if (peaks[i + 1].index - peaks[i].index < 2) {
            if (peaks[i].height >= peaks[i + 1].height) peaks.erase(peaks.begin() + (long)i + 1); else peaks.erase(peaks.begin() + (long)i);
        } else ++i;
    }
    for (size_t i = 0; peaks.size() >= 2 && i < peaks.size() - 1; i++) {      // determineTrends :728-756
        const Trend t = peaks[i].height < peaks[i + 1].height ? UP : peaks[i].height > peaks[i + 1].height ? DOWN : FLAG;
        peaks[i].right = t; peaks[i + 1].left = t;
    }
    if (peaks.size() == 1) peaks[0].main = true;                               // findMainPeakCandidates :758-799
    else
        for (size_t i = 0; i < peaks.size(); i++) {
            if (i == 0) peaks[i].main = peaks[i].right == DOWN;
            else if (i == peaks.size() - 1) peaks[i].main = peaks[i].left == UP;
            else peaks[i].main = peaks[i].left == UP && peaks[i].right == DOWN;
        }
    // setThresholdByValley :944-1039
    Valley low{0, 0.0, 0.0};
    double thr_pct = 0.0;
    int threshold = 0;
    std::vector<Peak> mains;                                                  // findFirstPriorityMainPeak :801-848
    for (const Peak &p : peaks) if (p.main) mains.push_back(p);
    if (mains.empty()) throw std::runtime_error("No main peaks found");
    size_t main_index;
    if (mains.size() == 1) main_index = mains[0].index;
    else {
        std::sort(mains.begin(), mains.end(), [](const Peak &a, const Peak &b) { return a.height > b.height; });
        main_index = mains[0].index > mains[1].index ? mains[0].index : mains[1].index;
    }
    bool found_sec = false;                                                   // findSecondaryPeak :850-909
    size_t sec_index = (size_t)-1;
    if (peaks.front().index != main_index) {
        size_t it = 0;
        while (peaks[it].index != main_index) it++;
        it--;
        if (it == 0) { sec_index = peaks[0].index; found_sec = true; }
        else {
            for (; it != 0; it--)
                if (peaks[it].left == DOWN && peaks[it].right == UP) { sec_index = peaks[it].index; found_sec = true; break; }
            if (!found_sec) { sec_index = peaks[0].index; found_sec = true; }
        }
    }
    if (found_sec) {
        const Peak sec = get_peak(peaks, sec_index, 0), next = get_peak(peaks, sec_index, 1);
        bool found = lowest_valley(h, sec.index, next.index, low);
        if (found) { thr_pct = low.percentage; threshold = (int)low.index; }
        if (thr_pct >= 0.3 || !found) {
            low = Valley{0, 0.0, 0.0}; thr_pct = 0.0; threshold = 0;
            if (sec.index != peaks[0].index) {
                const Peak pre = get_peak(peaks, sec.index, -1);
                if (lowest_valley(h, pre.index, sec.index, low)) { thr_pct = low.percentage; threshold = (int)low.index; }
            }
        }
    }
    if (low.height > max_height * 0.7) { thr_pct = 0.0; threshold = 0; }
    if (thr_pct >= 0.3) { thr_pct = 0.0; threshold = 0; }
    return threshold;
}

struct Box { double median = 0, q1 = 0, q3 = 0, iqr = 0, lo = 0, hi = 0; size_t n = 0; };

Box box_plot(std::vector<Feature> &feat) {                                    // statisticPurityData :281-343
    Box b;
    b.n = feat.size();
    if (b.n == 0) throw std::runtime_error("the data size is 0");
    std::sort(feat.begin(), feat.end(), [](const Feature &a, const Feature &c) { return a.ratio < c.ratio; });
    auto pct = [&](double p) -> double {
        const double pos = p * (b.n - 1);
        const size_t idx = (size_t)pos;
        const double frac = pos - idx;
        if (idx + 1 >= b.n) return feat[b.n - 1].ratio;
        return feat[idx].ratio * (1.0 - frac) + feat[idx + 1].ratio * frac;
    };
    b.q1 = pct(0.25); b.median = pct(0.5); b.q3 = pct(0.75);
    b.iqr = b.q3 - b.q1;
    b.lo = std::max(0.0, b.q1 - 1.5 * b.iqr);
    b.hi = b.q3 + 1.5 * b.iqr;
    return b;
}

}  // namespace

extern "C" int lps_estimate_purity(const lps_purity_input *in, lps_purity_result *out) {
    if (!in || !out || in->n < 0) return LPS_E_ARG;
    memset(out, 0, sizeof(*out));
    if (in->n && (!in->tumor_germline_imbalance || !in->normal_germline_imbalance || !in->normal_pct_germline_hp || !in->normal_h1 || !in->normal_h2))
        return LPS_E_ARG;
    if (in->used) memset(in->used, 0, (size_t)in->n);
    try {
        std::vector<Feature> feat;
        const double nor_min = (double)0.0f, tum_min = (double)0.0f, nor_max = (double)0.7f, pct_max = (double)0.7f;   // float thresholds (.h:280-283)
        for (int i = 0; i < in->n; i++) {
            const double nr = in->normal_germline_imbalance[i], tr = in->tumor_germline_imbalance[i];
            const int cnt = in->normal_h1[i] + in->normal_h2[i];
            if (nr == nor_min) out->filtered_normal_imbalance_zero++;
            else if (tr == tum_min) out->filtered_tumor_imbalance_zero++;
            else if (nr >= nor_max) out->filtered_normal_imbalance_high++;
            else if (cnt <= 5) out->filtered_normal_read_count++;
            else if (in->normal_pct_germline_hp[i] <= pct_max) out->filtered_pct_germline_hp++;
            else feat.push_back(Feature{i, tr, cnt});
        }
        out->n_after_lcvf = (int32_t)feat.size();
        if (feat.empty()) throw std::runtime_error("empty feature vector");
        int thr = 0;
        try { thr = valley_threshold(feat); } catch (const std::exception &) { thr = 0; }   // :227-231: a failure means threshold 0
        out->read_count_threshold = thr;
        {
            std::vector<Feature> kept;
            for (const Feature &f : feat) if (!(f.count < thr)) kept.push_back(f); else out->filtered_valley++;
            feat.swap(kept);
        }
        Box b = box_plot(feat);
        {
            std::vector<Feature> kept;
            for (const Feature &f : feat) if (f.ratio < b.lo || f.ratio > b.hi) out->filtered_outliers++; else kept.push_back(f);
            feat.swap(kept);
        }
        b = box_plot(feat);
        out->median = b.median; out->q1 = b.q1; out->q3 = b.q3; out->iqr = b.iqr; out->lower_whisker = b.lo; out->upper_whisker = b.hi;
        out->n_used = (int32_t)feat.size();
        for (const Feature &f : feat) out->n_outliers_left += (f.ratio < b.lo || f.ratio > b.hi);   // statisticPurityData :330-335
        if (in->used) for (const Feature &f : feat) in->used[f.idx] = 1;
        const double m = b.median, q = b.iqr;
        double purity = -3.3454 * m + 14.7747 * q + 4.0344 * m * m + -13.7777 * m * q + -5.2434 * q * q + 0.3058;
        if (purity > 1.0) purity = 1.0;
        else if (purity < 0.0) throw std::runtime_error("purity exceeds the model's estimation range");
        out->purity = purity;
        out->ok = 1;
    } catch (const std::exception &) {
        out->purity = 0.0;      // :78-82: any failure sets the purity to 0.0
        out->ok = 0;
    }
    return LPS_OK;
}
