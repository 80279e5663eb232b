// gsr_cull.cuh — the conservative work-skipping tests of the blend kernels and the per-Gaussian thresholds they use.
//
// These only SKIP work: a (sub-tile, splat) pair may be dropped only if no pixel of the sub-tile could pass the
// reference's own per-pixel tests (power <= 0 and alpha >= 1/255, CR/forward.cu:337-345), so they never change which
// (pixel, splat) pairs contribute.  GSR_HD so that tests/hostcheck can hammer that property on the CPU.
#pragma once
#include "gsr_math.cuh"

#if defined(__CUDA_ARCH__)
#define GSR_FAST_DIV(a, b) __fdividef((a), (b))
#else
#define GSR_FAST_DIV(a, b) ((a) / (b))
#endif

namespace gsr {

// Conservative squared radius (pixels^2) outside of which the splat cannot reach alpha >= 1/255:
//   alpha = o*exp(power) >= 1/255  =>  power >= -ln(255 o),   power <= -|d|^2 / (2 lam_max)
//   =>  |d|^2 <= 2 lam_max ln(255 o).
// 3% + 0.5 px^2 slack covers the rounding of the conic and of the power evaluation; for very large
// splats (lam_max > 1000 px^2) the relative error of the evaluated power is no longer negligible
// against that slack, so they are never culled (+inf).
GSR_HD float cull_radius2(float lam_max, float opacity)
{
    if (!(lam_max <= 1000.f)) return INFINITY;
    const float L = fmaxf(logf(255.f * opacity), 0.f);
    return 2.06f * lam_max * L + 0.5f;   // NaN opacity -> NaN -> never culled (tests are !(d2 > r2))
}
// Threshold for the exact ellipse-vs-sub-tile test of the forward blend: the splat can only contribute
// where q(d) = -power(d) <= ln(255 o); 3% + 0.02 slack covers the rounding of the conic and of the
// power evaluation inside the circle above (|d|^2 <= 2.06*1000*5.6, conic entries <= 1/0.3).  The low
// 3 mantissa bits are overwritten with the SH clamp flags by the caller, hence the extra 1e-5.
GSR_HD float cull_power(float lam_max, float opacity)
{
    if (!(lam_max <= 1000.f)) return INFINITY;
    const float L = fmaxf(logf(255.f * opacity), 0.f);
    return (1.03f * L + 0.02f) * 1.00001f;
}

// Conservative test: can a splat centred at (x,y) with cull radius^2 rc2 touch the sub-tile whose pixel centres span
// [sx0, sx1] x [sy0, sy1]?  Written as !(d2 > rc2) so that a NaN radius never culls.
GSR_HD bool subtile_hit(float sx0, float sx1, float sy0, float sy1, float x, float y, float rc2)
{
    const float dx = fmaxf(0.f, fmaxf(sx0 - x, x - sx1));
    const float dy = fmaxf(0.f, fmaxf(sy0 - y, y - sy1));
    return !(dx * dx + dy * dy > rc2);
}

// Exact (up to the slack folded into `lim`) test: does the ellipse {q(d) <= lim}, q(d) = 0.5 d^T Q d, reach
// the sub-tile?  The minimum of the convex q over the rectangle is 0 if the centre is inside, else it
// lies on the (at most two) edges facing the centre, where q is a 1-D quadratic with a closed-form
// minimiser.  Written so that NaNs never cull.
GSR_HD bool subtile_hit_ellipse(float sx0, float sx1, float sy0, float sy1, float cx, float cy, float A, float B, float C,
                                float lim)
{
    const float lx = sx0 - cx, hx = sx1 - cx;   // rect in splat-centred coordinates
    const float ly = sy0 - cy, hy = sy1 - cy;
    const float ex = lx > 0.f ? lx : (hx < 0.f ? hx : 0.f);   // offset to the facing vertical edge (0: inside in x)
    const float ey = ly > 0.f ? ly : (hy < 0.f ? hy : 0.f);
    if (ex == 0.f && ey == 0.f) return true;
    float qmin = INFINITY;
    if (ex != 0.f) {
        const float dy = fminf(hy, fmaxf(ly, GSR_FAST_DIV(-B * ex, C)));
        qmin = 0.5f * (A * ex * ex + C * dy * dy) + B * ex * dy;
    }
    if (ey != 0.f) {
        const float dx = fminf(hx, fmaxf(lx, GSR_FAST_DIV(-B * ey, A)));
        qmin = fminf(qmin, 0.5f * (A * dx * dx + C * ey * ey) + B * dx * ey);
    }
    return !(qmin > lim);
}

}  // namespace gsr
