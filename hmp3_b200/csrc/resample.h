// Host-side sample-rate converter of the CMp3Enc mirror handle: what Csrc does between MP3_audio_encode's PCM and the
// encoder when source and encode rate differ by something other than 1:1 or 1:2 (Csrc cases 2-4; hmp3/src/srcc.cpp:
// 82-206 filter design, 208-330 rate factoring, 334-393 + 495-640 coefficient generation; hmp3/src/srccf.cpp: the
// filters).  Restated, not copied: one class, the three channel layouts (mono, two channels, two channels mixed down)
// as one template parameter.  Every float expression keeps the reference's operand types and order, because the
// converted samples feed a bit-exact encoder (the tests compare them bit for bit with the reference's, call by call).
//   case 2: target > source: linear interpolation, n phases
//   case 3: target < source, n * ntaps <= 780 coefficients: n windowed low-pass filters of ntaps taps (polyphase FIR)
//   case 4: anything else: stage 1 interpolates up by (i+1)/i into a small buffer, stage 2 is case 3 on that buffer
// One call produces 1152 sample frames at the encode rate and says how many source frames it used up.
#pragma once
#include <math.h>
#include <string.h>

namespace hmp3 {

class Resampler {
public:
    enum Layout { MONO = 0, DUAL = 1, TO_MONO = 2 };
    int ncase = 0;   // 0 none, 1 = 1:2, 2..4 as above
    int minbuf = 0;  // source frames a call may look at (srcc.cpp:185-190)

    // Returns minbuf, or 0 if the pair of rates cannot be handled (coefficient tables too small, no factoring).
    int init(int source, int target) {
        memset(&s_, 0, sizeof(s_));  // (srcc.cpp:738: identical streams on a second run)
        int src2 = source;           // rate between the two stages (case 4)
        const int taps_all = taps_for(source, target);
        const int n_all = target / gcd_like(source, target);
        if (source == target) ncase = 0;
        else if (2 * source == target) ncase = 1;
        else if (source < target) ncase = 2;
        else if (n_all * taps_all <= 780) ncase = 3;
        else ncase = 4;
        int source1 = source, target1 = source;
        if (ncase == 4) {
            src2 = split_rate(source, target);
            if (src2 <= 0) return 0;
            target1 = src2;
        }
        s_.ntaps1 = taps_for(source1, target1);
        s_.n1 = target1 / gcd_like(source1, target1);
        s_.k1 = source1 / target1;
        s_.m1 = (s_.n1 * source1 - target1 * s_.n1 * s_.k1) / target1;
        s_.totcoef1 = s_.ntaps1 * s_.n1;
        int cut1 = (int)(0.90 * s_.ntaps1 * target1 / source1 + 0.50);
        if (cut1 > s_.ntaps1) cut1 = s_.ntaps1;
        s_.ntaps = taps_for(src2, target);
        s_.n = target / gcd_like(src2, target);
        s_.k = src2 / target;
        s_.m = (s_.n * src2 - target * s_.n * s_.k) / target;
        s_.totcoef = s_.ntaps * s_.n;
        int cut = (int)(0.90 * s_.ntaps * target / src2 + 0.50);
        if (cut > s_.ntaps) cut = s_.ntaps;
        s_.am = s_.n;
        s_.am1 = s_.n1;
        minbuf = (int)(1152.0 * source / target + (s_.ntaps - 1) + 1);
        if (ncase == 4) minbuf += (128 + 4);
        if (s_.totcoef1 > (int)(sizeof(s_.coef1) / sizeof(float))) return 0;
        if (s_.totcoef > (int)(sizeof(s_.coef) / sizeof(float))) return 0;
        make_filters(s_.coef1, s_.ntaps1, cut1, s_.n1, s_.m1);
        make_filters(s_.coef, s_.ntaps, cut, s_.n, s_.m);
        return minbuf;
    }

    // x: source frames as float (interleaved when the layout has two channels), y: 1152 frames out (interleaved for
    // DUAL).  Returns the source frames consumed.  Cases 2-4 only.
    int run(Layout layout, const float *x, float *y) {
        switch (layout) {
        case MONO: return run_<MONO>(x, y);
        case DUAL: return run_<DUAL>(x, y);
        default: return run_<TO_MONO>(x, y);
        }
    }

    // Source frames the next call of run() reads, which for case 2 can be one or two MORE than minbuf: the reference
    // interpolates between frame k and frame k + 1 up to the last output sample (srccf.cpp:102-130, 280-312, 496-528)
    // and so looks past the buffering figure its own init returns; a caller of the handle stages this many.
    int reach(Layout layout) const {
        if (ncase != 2) return minbuf;
        int am = s_.am, k = 0, top = 1;
        for (int i = 0; i < 1152; i++) {
            if (layout != TO_MONO && k + 1 > top) top = k + 1;
            am -= s_.m;
            if (am <= 0) {
                am += s_.n;
                k++;
                if (layout == TO_MONO && k + 1 > top) top = k + 1;
            }
        }
        return top + 1 > minbuf ? top + 1 : minbuf;
    }

private:
    struct State {
        int nbuf, kbuf;                              // stage-1 output buffer: fill, read position
        int ntaps1, n1, k1, m1, totcoef1, am1, ic1;  // stage 1 (case 4)
        float coef1[21];
        int ntaps, n, k, m, totcoef, am, ic;         // the filter (stage 2 of case 4)
        float coef[1280];
        float buf[128 + 64], buf2[128 + 64];
    } s_;

    // ---- rate arithmetic (srcc.cpp:82-104, 208-330)
    static int taps_for(int source, int target) {
        int t = (12 * source + target / 2) / target;
        if (t > 48) t = 48;
        if (t < 1) t = 1;
        t = (t & (~1)) | 1;
        if (source <= target) t = 1;
        return t;
    }
    static int gcd_like(int s, int t) {  // product of the common factors, found the reference's way
        int cf = 1;
        for (int i = 2; i <= t; i++) {
            if (s != i * (s / i) || t != i * (t / i)) continue;
            cf = i * cf;
            s = s / i;
            t = t / i;
            i = 1;
        }
        return cf;
    }
    static int split_rate(int source, int target) {  // intermediate rate of the two-stage conversion, 0 = none
        if (source <= target) return source;
        const int cf = gcd_like(source, target);
        const int s = source / cf, t = target / cf;
        int fs1 = 0, ft1 = 0;
        for (int i = 7; i < t; i++) {
            if (s != i * (s / i)) continue;
            if (t != (i + 1) * (t / (i + 1))) continue;
            fs1 = i;
            ft1 = i + 1;
            const int ft2 = t / ft1;
            const int mid = ft1 * source / fs1;
            if (taps_for(mid, target) * ft2 <= 780) break;
        }
        if (fs1 == 0) return 0;
        return ft1 * source / fs1;
    }

    // ---- coefficient generation (srcc.cpp:334-393, 495-530, 573-640): `nfilters` phase filters of `ntaps` taps, in the
    // order the filter walks them (phase advances by m modulo nfilters)
    static void lowpass(float *b, int N, int n, float alpha) {
        const double x = (N - 1) / 2.0 + alpha;
        const double pi = 4.0 * atan(1.0);
        const double t = pi / (2 * N);
        const double scale = 1.0 / N;
        for (int p = 0; p < N; p++) b[p] = 0.0f;
        for (int p = 0; p < N; p++) {
            const double wp = p == 0 ? 1.0 : 2.0;
            for (int k = 0; k < n; k++) b[p] += (float)(scale * wp * cos(t * x * (2 * k + 1)) * cos(t * p * (2 * k + 1)));
        }
    }
    static void raised_window(float *v, int n) {
        const double pi = 4.0 * atan(1.0);
        const double t = 2.0 * pi / n;
        for (int i = 0; i < n; i++) {
            double w = 0.5 * (1.0 - cos((i + 0.5) * t));
            w = .5 + .5 * w;
            v[i] = (float)(w * v[i]);
        }
    }
    static void unit_sum(float *v, int n) {
        float sum = 0.0f;
        for (int i = 0; i < n; i++) sum += v[i];
        for (int i = 0; i < n; i++) v[i] = v[i] / sum;
    }
    static void make_filters(float *a, int ntaps, int ncutoff, int nfilters, int m) {
        int am = 0;
        for (int i = 0; i < nfilters; i++) {
            float alpha = ((float)am) / nfilters;
            if (ntaps == 1) a[0] = alpha;  // used as y = f0 + alpha * (f1 - f0)
            else if (ntaps == 2) {
                a[0] = 1.0f - alpha;
                a[1] = alpha;
            } else {
                alpha = alpha + 0.5f / nfilters - 0.5f;  // filter centre = 0
                lowpass(a, ntaps, ncutoff, alpha);
                raised_window(a, ntaps);
                unit_sum(a, ntaps);
            }
            am += m;
            if (am >= nfilters) am = am - nfilters;
            a += ntaps;
        }
    }

    // ---- the filters (srccf.cpp:100-256 mono, 278-456 two channels, 494-640 mixed down)
    // One source frame of the layout as the reference reads it.
    template <int L>
    static float mixed(const float *x, int k) {  // TO_MONO: (left + right) * 0.5 with the reference's promotion to double
        return (float)((x[2 * k] + x[2 * k + 1]) * 0.5);
    }
    void step_phase() {
        s_.ic++;
        if (s_.ic >= s_.totcoef) s_.ic = 0;
    }
    bool step_source() {  // true when the source position moves on by one more frame
        s_.am -= s_.m;
        if (s_.am <= 0) {
            s_.am += s_.n;
            return true;
        }
        return false;
    }

    template <int L>
    int interpolate(const float *x, float *y) {  // case 2
        int k = 0;
        if (L == TO_MONO) {
            float a = (float)((x[0] + x[1]) * 0.5);
            float b = (float)(((x[2] + x[3]) * 0.5) - a);
            for (int i = 0; i < 1152; i++) {
                y[i] = (float)(a + s_.coef[s_.ic] * b);
                step_phase();
                if (step_source()) {
                    k++;
                    a = a + b;
                    b = (float)(((x[2 * (k + 1)] + x[2 * (k + 1) + 1]) * 0.5) - a);
                }
            }
            return k;
        }
        for (int i = 0; i < 1152; i++) {
            if (L == MONO) y[i] = (float)((float)x[k] + s_.coef[s_.ic] * ((float)x[k + 1] - (float)x[k]));
            else {
                y[2 * i] = (float)((float)x[2 * k] + s_.coef[s_.ic] * ((float)x[2 * (k + 1)] - (float)x[2 * k]));
                y[2 * i + 1] = (float)((float)x[2 * k + 1] + s_.coef[s_.ic] * ((float)x[2 * (k + 1) + 1] - (float)x[2 * k + 1]));
            }
            step_phase();
            if (step_source()) k++;
        }
        return k;
    }

    template <int L>
    int decimate(const float *x, float *y) {  // case 3
        int k = 0;
        for (int i = 0; i < 1152; i++) {
            float u = 0.0f, v = 0.0f;
            for (int j = 0; j < s_.ntaps; j++) {
                if (L == MONO) u += s_.coef[s_.ic++] * x[k + j];
                else if (L == DUAL) {
                    u += s_.coef[s_.ic] * x[2 * (k + j)];
                    v += s_.coef[s_.ic++] * x[2 * (k + j) + 1];
                } else u += s_.coef[s_.ic++] * ((x[2 * (k + j)] + x[2 * (k + j) + 1]) * 0.5);
            }
            if (L == DUAL) {
                y[2 * i] = u;
                y[2 * i + 1] = v;
            } else y[i] = u;
            if (s_.ic >= s_.totcoef) s_.ic = 0;
            k += s_.k;
            if (step_source()) k++;
        }
        return k;
    }

    template <int L>
    int first_stage(const float *x) {  // 128 more interpolated frames into the buffer; returns the source frames used
        s_.nbuf -= s_.kbuf;
        if (s_.nbuf > 0) {
            memmove(s_.buf, s_.buf + s_.kbuf, sizeof(float) * s_.nbuf);
            if (L == DUAL) memmove(s_.buf2, s_.buf2 + s_.kbuf, sizeof(float) * s_.nbuf);
        }
        s_.kbuf = 0;
        int j = 0;
        float a = 0.0f, b = 0.0f;
        if (L == TO_MONO) {
            a = mixed<L>(x, 0);
            b = mixed<L>(x, 1);
        }
        for (int i = 0; i < 128; i++) {
            if (L == MONO) s_.buf[s_.nbuf++] = (float)x[j] + s_.coef1[s_.ic1] * ((float)x[j + 1] - (float)x[j]);
            else if (L == DUAL) {
                s_.buf[s_.nbuf] = (float)x[2 * j] + s_.coef1[s_.ic1] * ((float)x[2 * (j + 1)] - (float)x[2 * j]);
                s_.buf2[s_.nbuf++] = (float)x[2 * j + 1] + s_.coef1[s_.ic1] * ((float)x[2 * (j + 1) + 1] - (float)x[2 * j + 1]);
            } else s_.buf[s_.nbuf++] = a + s_.coef1[s_.ic1] * (b - a);
            s_.ic1++;
            if (s_.ic1 >= s_.totcoef1) s_.ic1 = 0;
            s_.am1 -= s_.m1;
            if (s_.am1 <= 0) {
                s_.am1 += s_.n1;
                j++;
                if (L == TO_MONO) {
                    a = b;
                    b = mixed<L>(x, j + 1);
                }
            }
        }
        return j;
    }

    template <int L>
    int two_stage(const float *x, float *y) {  // case 4
        const int fw = (L == MONO) ? 1 : 2;  // floats per source frame
        int thres = s_.nbuf - s_.ntaps;
        int k0 = 0;
        for (int i = 0; i < 1152; i++) {
            if (s_.kbuf > thres) {
                k0 += first_stage<L>(x + fw * k0);
                thres = s_.nbuf - s_.ntaps;
            }
            float u = 0.0f, v = 0.0f;
            for (int j = 0; j < s_.ntaps; j++) {
                if (L == DUAL) {
                    u += s_.coef[s_.ic] * s_.buf[s_.kbuf + j];
                    v += s_.coef[s_.ic++] * s_.buf2[s_.kbuf + j];
                } else u += s_.coef[s_.ic++] * s_.buf[s_.kbuf + j];
            }
            if (L == DUAL) {
                y[2 * i] = u;
                y[2 * i + 1] = v;
            } else y[i] = u;
            if (s_.ic >= s_.totcoef) s_.ic = 0;
            s_.kbuf += s_.k;
            if (step_source()) s_.kbuf++;
        }
        return k0;
    }

    template <int L>
    int run_(const float *x, float *y) {
        if (ncase == 2) return interpolate<L>(x, y);
        if (ncase == 3) return decimate<L>(x, y);
        return two_stage<L>(x, y);
    }
};

}  // namespace hmp3
