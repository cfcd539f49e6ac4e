// tight_bbox.h -- provably empty border columns / rows of a triangle's pixel bounding box.
//
// The reference walks every pixel of  [trunc(min), ceil(max)]  (bounding_box, drawing.cpp:77-93) and rejects
// most of them with the barycentric test (drawing.cpp:41-49,111).  A triangle smaller than a pixel therefore
// costs 4-9 pixel tests although it usually contains no sample point at all: the first column lies left of
// every vertex, the last one right of every vertex, likewise the rows.  This header decides, per border
// column / row, whether the reference's own fp32 test is CERTAIN to reject every pixel of it; only then is it
// dropped, so the set of accepted pixels -- and with it every output bit -- is unchanged.
//
// Notation: vertices v_k = (x_k, y_k), pixel p, u = 2^-24.  E_k(p) = exact value of the reference's edge function
// opposite v_k, e_k(p) = its fp32 evaluation  fl(fl(fl(b.x-a.x)*fl(p.y-a.y)) - fl(fl(b.y-a.y)*fl(p.x-a.x)))
// (drawing.cpp:36-39), A = E_2(v_2) = exact doubled area, `area` = its fp32 evaluation (drawing.cpp:46),
// w = max x_k - min x_k, h = max y_k - min y_k.
//
//  (1) Rounding.  Two rounded differences, one rounded product per term and one rounded subtraction give
//        |e_k(p) - E_k(p)| <= 4.000001 u (|b.x-a.x||p.y-a.y| + |b.y-a.y||p.x-a.x|) + 2^-148
//                          <= 4.000001 u (w DY + h DX) + 2^-148 =: errE
//      for every pixel p of the bbox, where DX (DY) bounds |p.x - x_k| (|p.y - y_k|) over the bbox and the
//      vertices.  In the same way |area - A| <= 8.000002 u w h + 2^-148 =: errA; if |area| > errA, A has the sign
//      of `area` and |A| >= |area| - errA.
//  (2) Geometry (exact).  Barycentric identities: sum_k E_k(p) = A and sum_k E_k(p) (x_k - p.x) = 0.  Let p lie
//      left of every vertex by d = min x_k - p.x > 0, so every m_k = x_k - p.x is in [d, d + w].  Writing
//      E'_k = sign(A) E_k and N for the set of negative E'_k:  d (|A| + sum_N |E'_k|) <= sum_{not N} E'_k m_k
//      = sum_N |E'_k| m_k <= (d + w) sum_N |E'_k|,  hence  sum_N |E'_k| >= d |A| / w,  and since at most two of
//      the three can be negative,  min_k E'_k <= - d |A| / (2 w).  The same holds right of every vertex, and
//      above / below with h in place of w.
//  (3) The reference accepts a pixel only if every quotient e_k / area >= 0, which requires
//      sign(area) e_k >= -2^-22 (the quotient of a negative numerator can only be "-0 >= 0" by underflow:
//      kernels.cuh, candidate()).  With (1) and (2):  sign(area) e_k(p) <= -d (|area| - errA) / (2 w) + errE for
//      the minimising k, so the pixel is certainly rejected when
//            d (|area| - errA) > 2 w (errE + 2^-22).                                                     (*)
//
// The code evaluates (*) in fp32 with twice the constants (8 u, 16 u, 2^-21) and absolute floors, which covers
// the few roundings of the bound itself with a wide margin; NaN, infinities and |area| <= errA are
// screened first (nothing is dropped).  Only the outermost column / row of each side is examined: the bbox
// extends less than one pixel beyond the vertices unless it was clamped to the image, and a clamped side is a
// single column / row.  tests/tight_rule_check.c replays the reference's literal test over billions of bbox
// pixels (sub-pixel triangles, slivers, near-degenerate and off-screen ones, huge and denormal coordinates) and
// checks that no dropped pixel is accepted and that each one fails the candidate test with the predicted margin.
#pragma once

#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define RAST_TIGHT_HD __host__ __device__ __forceinline__
#else
#define RAST_TIGHT_HD static inline
#endif

// Shrinks the inclusive pixel rectangle [*x0,*x1] x [*y0,*y1] (the reference's bbox, possibly clamped to the image
// and to a band of rows) of the triangle (ax,ay) (bx,by) (cx,cy).  area_abs = |fp32 area| as the reference computes
// it (drawing.cpp:46).  Returns 0 when no pixel is left.  Rectangle coordinates must be below 2^24.
RAST_TIGHT_HD int rast_tight_bbox(float ax, float ay, float bx, float by, float cx, float cy, float area_abs,
                                  uint32_t *x0, uint32_t *y0, uint32_t *x1, uint32_t *y1) {
    if (!(area_abs > 0.f && area_abs < 3.0e38f)) return 1; // 0, inf, NaN: the literal path decides (candidate() is always true there)
    const float minx = fminf(fminf(ax, bx), cx), maxx = fmaxf(fmaxf(ax, bx), cx);
    const float miny = fminf(fminf(ay, by), cy), maxy = fmaxf(fmaxf(ay, by), cy);
    const float w = maxx - minx, h = maxy - miny;
    const float fx0 = (float)*x0, fx1 = (float)*x1, fy0 = (float)*y0, fy1 = (float)*y1;
    const float DX = fmaxf(fx1 - minx, maxx - fx0), DY = fmaxf(fy1 - miny, maxy - fy0); // >= |p - v_k| per axis over the bbox
    const float errE = (w * DY + h * DX) * 4.76837158203125e-07f /* 8u = 2^-21 */ + 7.17e-43f /* ~2^-140 */;
    const float errA = (w * h) * 9.5367431640625e-07f /* 16u = 2^-20 */ + 7.17e-43f;
    const float denom = area_abs - errA;
    if (!(denom > 0.f)) return 1; // the sign of the exact area is not certain (or NaN): nothing is dropped
    const float slack = errE + 4.76837158203125e-07f /* 2 * 2^-22 */;
    const float rx = (w + w) * slack + 1e-30f, ry = (h + h) * slack + 1e-30f;
    int32_t xs = (int32_t)*x0, xe = (int32_t)*x1, ys = (int32_t)*y0, ye = (int32_t)*y1;
    if ((minx - fx0) * denom > rx) ++xs; // first column left of every vertex by enough
    if ((fx1 - maxx) * denom > rx) --xe; // last column right of every vertex
    if ((miny - fy0) * denom > ry) ++ys;
    if ((fy1 - maxy) * denom > ry) --ye;
    if (xs > xe || ys > ye) return 0;
    *x0 = (uint32_t)xs; *x1 = (uint32_t)xe; *y0 = (uint32_t)ys; *y1 = (uint32_t)ye;
    return 1;
}
