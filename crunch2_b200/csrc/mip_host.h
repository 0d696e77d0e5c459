// mip_host.h -- host side of the resampler: filter kernels and contributor lists with the reference's arithmetic
// (crnlib/crn_resample_filters.cpp, Resampler::make_clist in crnlib/crn_resampler.cpp:119-420), the gamma tables of
// image_utils::resample_multithreaded (crnlib/crn_image_utils.cpp:686-712).  Plain C++ on purpose: these are a few
// thousand libm calls per level; the per-pixel work is in mip_kernels.cuh.
#pragma once
#include <math.h>
#include <vector>

namespace {

double mipf_sinc(double x)
{   // crn_resample_filters.cpp:182-191
    x = (x * 3.14159265358979323846);
    if ((x < 0.01f) && (x > -0.01f)) return 1.0f + x * x * (-1.0f / 6.0f + x * x * 1.0f / 120.0f);
    return sin(x) / x;
}
float mipf_clean(double t) { const float EPSILON = .0000125f; if (fabs(t) < EPSILON) return 0.0f; return static_cast<float>(t); }
double mipf_bessel0(double x)
{   // :322-342
    const double EPSILON_RATIO = 1E-16;
    double xh = 0.5 * x, sum = 1.0, pw = 1.0, ds = 1.0;
    int k = 0;
    while (ds > sum * EPSILON_RATIO) { ++k; pw = pw * (xh / k); ds = pw * pw; sum = sum + ds; }
    return sum;
}
double mipf_kaiser(double alpha, double half_width, double x) { const double ratio = (x / half_width); return mipf_bessel0(alpha * sqrt(1 - ratio * ratio)) / mipf_bessel0(alpha); }
float mip_filter_box(float t) { return ((t >= -0.5f) && (t < 0.5f)) ? 1.0f : 0.0f; }
float mip_filter_tent(float t) { if (t < 0.0f) t = -t; return t < 1.0f ? 1.0f - t : 0.0f; }
float mip_filter_lanczos4(float t) { if (t < 0.0f) t = -t; return t < 4.0f ? mipf_clean(mipf_sinc(t) * mipf_sinc(t / 4.0f)) : 0.0f; }
float mip_filter_mitchell(float t)
{   // mitchell(t, 1/3, 1/3), crn_resample_filters.cpp:143-167
    const float B = 1.0f / 3.0f, C = 1.0f / 3.0f;
    float tt = t * t;
    if (t < 0.0f) t = -t;
    if (t < 1.0f) { t = (((12.0f - 9.0f * B - 6.0f * C) * (t * tt)) + ((-18.0f + 12.0f * B + 6.0f * C) * tt) + (6.0f - 2.0f * B)); return (t / 6.0f); }
    else if (t < 2.0f) { t = (((-1.0f * B - 6.0f * C) * (t * tt)) + ((6.0f * B + 30.0f * C) * tt) + ((-12.0f * B - 48.0f * C) * t) + (8.0f * B + 24.0f * C)); return (t / 6.0f); }
    return 0.0f;
}
float mip_filter_kaiser(float t)
{   // :352-369
    if (t < 0.0f) t = -t;
    if (t < 3) {
        const float att = 40.0f;
        const float alpha = (float)(exp(log((double)0.58417 * (att - 20.96)) * 0.4) + 0.07886 * (att - 20.96));
        return (float)mipf_clean(mipf_sinc(t) * mipf_kaiser(alpha, 3, t));
    }
    return 0.0f;
}
struct MipFilter { const char* name; float (*func)(float); float support; };
// crn_mip_filter numbering (inc/crnlib.h:438-446)
const MipFilter g_mip_filters[5] = { { "box", mip_filter_box, 0.5f }, { "tent", mip_filter_tent, 1.0f }, { "lanczos4", mip_filter_lanczos4, 4.0f },
                                     { "mitchell", mip_filter_mitchell, 2.0f }, { "kaiser", mip_filter_kaiser, 3.0f } };

struct MipContribs { std::vector<uint32_t> off, pix; std::vector<float> wgt; };

int mip_posmod(int x, int y) { if (x >= 0) return (x < y) ? x : (x % y); int m = (-x) % y; return (m != 0) ? (y - m) : m; }
int mip_reflect(int j, int src_x, bool wrap)
{   // Resampler::reflect with BOUNDARY_WRAP / BOUNDARY_CLAMP (crn_resampler.cpp:65-114)
    if (j < 0) return wrap ? mip_posmod(j, src_x) : 0;
    if (j >= src_x) return wrap ? mip_posmod(j, src_x) : src_x - 1;
    return j;
}

// Resampler::make_clist(src_x, dst_x, boundary_op, Pfilter, filter_support, filter_scale, src_ofs = 0)
bool mip_make_clist(int src_x, int dst_x, bool wrap, const MipFilter& F, float filter_scale, MipContribs& out)
{
    out.off.assign(1, 0u); out.pix.clear(); out.wgt.clear();
    const float oo_filter_scale = 1.0f / filter_scale;
    const float NUDGE = 0.5f;
    const float xscale = dst_x / (float)src_x;
    const bool down = xscale < 1.0f;
    const float half_width = down ? (F.support / xscale) * filter_scale : F.support * filter_scale;
    std::vector<float> fv;
    for (int i = 0; i < dst_x; i++) {
        float center = ((float)i + NUDGE) / xscale;
        center -= NUDGE;
        center += 0.0f;
        const int left = (int)(float)floor(center - half_width), right = (int)(float)ceil(center + half_width);
        // the reference evaluates the filter twice per tap (once to normalise, once to weight); the value is the same both
        // times, so it is computed once and kept
        fv.resize((size_t)(right - left + 1));
        float total_weight = 0;
        for (int j = left; j <= right; j++) {
            fv[(size_t)(j - left)] = down ? F.func((center - (float)j) * xscale * oo_filter_scale) : F.func((center - (float)j) * oo_filter_scale);
            total_weight += fv[(size_t)(j - left)];
        }
        const float norm = static_cast<float>(1.0f / total_weight);
        total_weight = 0;
        int max_k = -1; float max_w = -1e+20f;
        const size_t first = out.pix.size();
        for (int j = left; j <= right; j++) {
            const float weight = fv[(size_t)(j - left)] * norm;
            if (weight == 0.0f) continue;
            const int n = mip_reflect(j, src_x, wrap);
            const int k = (int)(out.pix.size() - first);
            out.pix.push_back((uint32_t)(unsigned short)n); out.wgt.push_back(weight);
            total_weight += weight;
            if (weight > max_w) { max_w = weight; max_k = k; }
        }
        if (max_k == -1 || out.pix.size() == first) return false;
        if (total_weight != 1.0f) out.wgt[first + max_k] += 1.0f - total_weight;
        out.off.push_back((uint32_t)out.pix.size());
    }
    return true;
}

}  // namespace
