/*
 * tight_rule_check.c -- TEST ONLY.  Brute-force check of rasteriser_b200/csrc/tight_bbox.h against the reference's
 * literal per-pixel test.
 *
 * For random triangles in raster space (regimes below) the reference's bounding box is computed exactly as
 * bounding_box does (drawing.cpp:77-93: glm::min/max, ceil, clamp, unsigned cast), rast_tight_bbox() shrinks it,
 * and EVERY pixel the rule dropped is put through the reference's own test (barycentric, drawing.cpp:41-49:
 * three edge functions divided by the area; inside = all >= 0, drawing.cpp:111), restated here operation for
 * operation like oracle/oracle.c:392-398.  Two claims are checked per dropped pixel:
 *   weak   (what parity needs)  : the literal test rejects it;
 *   strong (what the proof says): min_k sign(area)*e_k < -2^-22, i.e. it is not even a candidate (kernels.cuh);
 * the smallest ratio  min_k sign(area)*e_k / -2^-22  seen is reported (the proof predicts > 2).
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off -o tight_rule_check tight_rule_check.c -lm      (no FMA contraction)
 * Usage: tight_rule_check <seed> <triangles>   -> one JSON line
 *        tight_rule_check --file tris.bin W H   (float32 [n][3][2] raster-space triangles of a real scene)
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../rasteriser_b200/csrc/tight_bbox.h"

static uint64_t rng_state;
static uint64_t rnd(void) { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static double uni(void) { return (double)(rnd() >> 11) * (1.0 / 9007199254740992.0); }            /* [0,1) */
static double range(double a, double b) { return a + (b - a) * uni(); }
static float bits_float(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }

/* glm 0.9.7 min/max and the x86-64 float -> unsigned cast, as in oracle/oracle.c:31-32,269 */
static inline float glm_min(float x, float y) { return x < y ? x : y; }
static inline float glm_max(float x, float y) { return x > y ? x : y; }
static inline uint32_t float_to_uint(float f) { return (uint32_t)(long long)f; }
/* drawing.cpp:36-39 */
static inline float edge(float px, float py, const float *a, const float *b) { return (b[0] - a[0]) * (py - a[1]) - (b[1] - a[1]) * (px - a[0]); }

typedef struct {
    uint64_t triangles, shrunk, emptied, bbox_pixels, dropped_pixels, checked_pixels, weak_violations, strong_violations;
    double weakest_ratio;
} stats;

static void check_pixel(const float v[3][2], float area, uint32_t x, uint32_t y, stats *st) {
    const float px = (float)x, py = (float)y;
    const float e0 = edge(px, py, v[1], v[2]), e1 = edge(px, py, v[2], v[0]), e2 = edge(px, py, v[0], v[1]);
    const float b0 = e0 / area, b1 = e1 / area, b2 = e2 / area;
    st->checked_pixels++;
    if (b0 >= 0.f && b1 >= 0.f && b2 >= 0.f) {
        if (st->weak_violations++ < 5)
            fprintf(stderr, "WEAK VIOLATION px (%u,%u) v (%a,%a) (%a,%a) (%a,%a) area %a\n", x, y, v[0][0], v[0][1], v[1][0], v[1][1], v[2][0], v[2][1], area);
    }
    const float s = area < 0.f ? -1.f : 1.f;
    const float m = fminf(fminf(s * e0, s * e1), s * e2);
    const int nan = (e0 != e0) || (e1 != e1) || (e2 != e2);
    if (nan || !(m < -2.384185791015625e-07f)) {
        if (st->strong_violations++ < 5)
            fprintf(stderr, "STRONG VIOLATION px (%u,%u) min %a v (%a,%a) (%a,%a) (%a,%a) area %a\n", x, y, m, v[0][0], v[0][1], v[1][0], v[1][1], v[2][0], v[2][1], area);
    } else {
        const double ratio = (double)m / -2.384185791015625e-07;
        if (ratio < st->weakest_ratio) st->weakest_ratio = ratio;
    }
}

static void one_triangle(float v[3][2], uint32_t W, uint32_t H, uint32_t band0, uint32_t band1, stats *st) {
    st->triangles++;
    /* bounding_box (drawing.cpp:77-93), as oracle/oracle.c:376-385 */
    const float brx = (float)(W - 1u), bry = (float)(H - 1u);
    const float minx = glm_min(glm_min(v[0][0], v[1][0]), v[2][0]), miny = glm_min(glm_min(v[0][1], v[1][1]), v[2][1]);
    const float maxx = ceilf(glm_max(glm_max(v[0][0], v[1][0]), v[2][0])), maxy = ceilf(glm_max(glm_max(v[0][1], v[1][1]), v[2][1]));
    uint32_t x0 = float_to_uint(glm_min(glm_max(minx, 0.f), brx)), y0 = float_to_uint(glm_min(glm_max(miny, 0.f), bry));
    uint32_t x1 = float_to_uint(glm_min(glm_max(maxx, 0.f), brx)), y1 = float_to_uint(glm_min(glm_max(maxy, 0.f), bry));
    /* band of rows (the device clamps the bbox to its band the same way, kernels.cuh bounding_box) */
    if (y0 < band0) y0 = band0;
    if (y1 >= band1) y1 = band1 - 1u;
    if (x1 < x0 || y1 < y0) return;
    st->bbox_pixels += (uint64_t)(x1 - x0 + 1u) * (y1 - y0 + 1u);

    const float area = edge(v[2][0], v[2][1], v[0], v[1]); /* drawing.cpp:46 */
    uint32_t tx0 = x0, ty0 = y0, tx1 = x1, ty1 = y1;
    const int left = rast_tight_bbox(v[0][0], v[0][1], v[1][0], v[1][1], v[2][0], v[2][1], fabsf(area), &tx0, &ty0, &tx1, &ty1);
    if (!left) { st->emptied++; tx0 = x1 + 1u; tx1 = x1; ty0 = y0; ty1 = y1; } /* everything dropped */
    else if (tx0 != x0 || ty0 != y0 || tx1 != x1 || ty1 != y1) st->shrunk++;
    if (left && (tx0 < x0 || tx1 > x1 || ty0 < y0 || ty1 > y1 || tx0 > tx1 || ty0 > ty1)) { st->weak_violations++; fprintf(stderr, "rule grew the rectangle\n"); return; }

    const uint64_t full = (uint64_t)(x1 - x0 + 1u) * (y1 - y0 + 1u), kept = left ? (uint64_t)(tx1 - tx0 + 1u) * (ty1 - ty0 + 1u) : 0u;
    st->dropped_pixels += full - kept;
    if (full == kept) return;
    if (full <= 20000u) {
        for (uint32_t y = y0; y <= y1; ++y)
            for (uint32_t x = x0; x <= x1; ++x)
                if (!left || x < tx0 || x > tx1 || y < ty0 || y > ty1) check_pixel((const float (*)[2])v, area, x, y, st);
    } else { /* large bbox: the ends and a random sample of each dropped column / row */
        for (int side = 0; side < 4; ++side) {
            const int dropped = !left || (side == 0 ? tx0 != x0 : side == 1 ? tx1 != x1 : side == 2 ? ty0 != y0 : ty1 != y1);
            if (!dropped) continue;
            for (int k = 0; k < 512; ++k) {
                uint32_t x, y;
                if (side < 2) { x = side == 0 ? x0 : x1; y = k == 0 ? y0 : k == 1 ? y1 : y0 + (uint32_t)(rnd() % (y1 - y0 + 1u)); }
                else { y = side == 2 ? y0 : y1; x = k == 0 ? x0 : k == 1 ? x1 : x0 + (uint32_t)(rnd() % (x1 - x0 + 1u)); }
                check_pixel((const float (*)[2])v, area, x, y, st);
            }
        }
    }
}

static float nudge(float f, int ulps) { for (int i = 0; i < abs(ulps); ++i) f = nextafterf(f, ulps > 0 ? INFINITY : -INFINITY); return f; }

/* --file tris.bin W H: float32 [n][3][2] raster-space triangles of a real scene (tests/test_tight_bbox_rule.py) */
static int run_file(const char *path, uint32_t W, uint32_t H) {
    FILE *f = fopen(path, "rb");
    if (!f) { perror(path); return 2; }
    stats st;
    memset(&st, 0, sizeof st);
    st.weakest_ratio = 1e300;
    rng_state = 0x9E3779B97F4A7C15ull;
    float v[3][2];
    while (fread(v, sizeof v, 1, f) == 1) one_triangle(v, W, H, 0, H, &st);
    fclose(f);
    printf("{\"triangles\": %llu, \"shrunk\": %llu, \"emptied\": %llu, \"bbox_pixels\": %llu, \"dropped_pixels\": %llu, \"checked_pixels\": %llu, "
           "\"weak_violations\": %llu, \"strong_violations\": %llu, \"weakest_ratio\": %.6g}\n",
           (unsigned long long)st.triangles, (unsigned long long)st.shrunk, (unsigned long long)st.emptied, (unsigned long long)st.bbox_pixels,
           (unsigned long long)st.dropped_pixels, (unsigned long long)st.checked_pixels, (unsigned long long)st.weak_violations,
           (unsigned long long)st.strong_violations, st.weakest_ratio);
    return (st.weak_violations || st.strong_violations) ? 1 : 0;
}

int main(int argc, char **argv) {
    if (argc == 5 && !strcmp(argv[1], "--file")) return run_file(argv[2], (uint32_t)atoi(argv[3]), (uint32_t)atoi(argv[4]));
    const uint64_t seed = argc > 1 ? strtoull(argv[1], NULL, 10) : 1u;
    const uint64_t n = argc > 2 ? strtoull(argv[2], NULL, 10) : 1000000u;
    rng_state = 0x9E3779B97F4A7C15ull ^ (seed * 0xD1B54A32D192ED03ull + 1u);
    for (int i = 0; i < 8; ++i) rnd();
    static const uint32_t sizes[] = {1, 2, 3, 7, 64, 640, 1920, 3840, 7680, 65535};
    stats st;
    memset(&st, 0, sizeof st);
    st.weakest_ratio = 1e300;
    for (uint64_t t = 0; t < n; ++t) {
        const uint32_t W = sizes[rnd() % 10], H = sizes[rnd() % 10];
        uint32_t band0 = 0, band1 = H;
        if (rnd() % 4 == 0) { band0 = (uint32_t)(rnd() % H); band1 = band0 + 1u + (uint32_t)(rnd() % (H - band0)); }
        float v[3][2];
        const int regime = (int)(rnd() % 12);
        const double cx = range(-3.0, (double)W + 3.0), cy = range(-3.0, (double)H + 3.0);
        switch (regime) {
        case 0: case 1: { /* sub-pixel to a few pixels, anywhere around the image */
            const double s = pow(10.0, range(-3.0, 0.7));
            for (int k = 0; k < 3; ++k) { v[k][0] = (float)(cx + s * range(-1, 1)); v[k][1] = (float)(cx * 0 + cy + s * range(-1, 1)); }
        } break;
        case 2: { /* slivers */
            const double th = range(0, 6.283185307), L = pow(10.0, range(-1.0, 2.2)), thick = pow(10.0, range(-7.0, -1.0)), al = uni();
            v[0][0] = (float)cx; v[0][1] = (float)cy;
            v[1][0] = (float)(cx + L * cos(th)); v[1][1] = (float)(cy + L * sin(th));
            v[2][0] = (float)(cx + al * L * cos(th) - thick * sin(th)); v[2][1] = (float)(cy + al * L * sin(th) + thick * cos(th));
        } break;
        case 3: { /* nearly collinear: third vertex a few ulps off the line through the other two */
            const double s = pow(10.0, range(-2.0, 1.5)), al = range(-0.5, 1.5);
            v[0][0] = (float)(cx + s * range(-1, 1)); v[0][1] = (float)(cy + s * range(-1, 1));
            v[1][0] = (float)(cx + s * range(-1, 1)); v[1][1] = (float)(cy + s * range(-1, 1));
            v[2][0] = nudge((float)(v[0][0] + al * ((double)v[1][0] - v[0][0])), (int)(rnd() % 9) - 4);
            v[2][1] = nudge((float)(v[0][1] + al * ((double)v[1][1] - v[0][1])), (int)(rnd() % 9) - 4);
        } break;
        case 4: { /* lattice: vertices on multiples of 1/8 -- integer extents, edges through sample points */
            const double ix = floor(cx), iy = floor(cy);
            for (int k = 0; k < 3; ++k) { v[k][0] = (float)(ix + (double)((int)(rnd() % 33) - 16) / 8.0); v[k][1] = (float)(iy + (double)((int)(rnd() % 33) - 16) / 8.0); }
        } break;
        case 5: { /* far off-screen / huge coordinates */
            const double sc = pow(10.0, range(3.0, 30.0)), s = sc * pow(10.0, range(-8.0, 0.0));
            const double ox = sc * range(-1, 1), oy = sc * range(-1, 1);
            for (int k = 0; k < 3; ++k) { v[k][0] = (float)(ox + s * range(-1, 1)); v[k][1] = (float)(oy + s * range(-1, 1)); }
        } break;
        case 6: { /* tiny and denormal coordinates next to the origin */
            const double s = pow(10.0, range(-45.0, -18.0));
            for (int k = 0; k < 3; ++k) { v[k][0] = (float)(s * range(-1, 1)); v[k][1] = (float)(s * range(-1, 1)); }
        } break;
        case 7: { /* the size range the setup thread walks itself (bbox up to 64 pixels) and the overflow fallback */
            const double s = pow(10.0, range(0.0, 2.0));
            for (int k = 0; k < 3; ++k) { v[k][0] = (float)(cx + s * range(-1, 1)); v[k][1] = (float)(cy + s * range(-1, 1)); }
        } break;
        case 8: { /* arbitrary bit patterns (NaN, infinities, denormals included) */
            for (int k = 0; k < 3; ++k) { v[k][0] = bits_float((uint32_t)rnd()); v[k][1] = bits_float((uint32_t)rnd()); }
        } break;
        case 9: { /* around the origin, where nearby floats do not subtract exactly */
            const double s = pow(10.0, range(-6.0, 0.5));
            for (int k = 0; k < 3; ++k) { v[k][0] = (float)(range(-1.5, 2.5) + s * range(-1, 1)); v[k][1] = (float)(range(-1.5, 2.5) + s * range(-1, 1)); }
        } break;
        case 10: { /* extents a few ulps away from integers: the dropped column is only just outside */
            const double s = pow(10.0, range(-1.0, 0.8));
            for (int k = 0; k < 3; ++k) { v[k][0] = (float)(cx + s * range(-1, 1)); v[k][1] = (float)(cy + s * range(-1, 1)); }
            const int k = (int)(rnd() % 3);
            v[k][0] = nudge(floorf(v[k][0]) + (float)(rnd() % 2), (int)(rnd() % 7) - 3);
            v[k][1] = nudge(floorf(v[k][1]) + (float)(rnd() % 2), (int)(rnd() % 7) - 3);
        } break;
        default: { /* small triangle with one vertex exactly on a sample point */
            const double s = pow(10.0, range(-2.0, 0.5));
            for (int k = 0; k < 3; ++k) { v[k][0] = (float)(cx + s * range(-1, 1)); v[k][1] = (float)(cy + s * range(-1, 1)); }
            v[0][0] = floorf(v[0][0]); v[0][1] = floorf(v[0][1]);
        } break;
        }
        one_triangle(v, W, H, band0, band1, &st);
    }
    printf("{\"seed\": %llu, \"triangles\": %llu, \"shrunk\": %llu, \"emptied\": %llu, \"bbox_pixels\": %llu, \"dropped_pixels\": %llu, \"checked_pixels\": %llu, "
           "\"weak_violations\": %llu, \"strong_violations\": %llu, \"weakest_ratio\": %.6g}\n",
           (unsigned long long)seed, (unsigned long long)st.triangles, (unsigned long long)st.shrunk, (unsigned long long)st.emptied,
           (unsigned long long)st.bbox_pixels, (unsigned long long)st.dropped_pixels, (unsigned long long)st.checked_pixels,
           (unsigned long long)st.weak_violations, (unsigned long long)st.strong_violations, st.weakest_ratio);
    return (st.weak_violations || st.strong_violations) ? 1 : 0;
}
