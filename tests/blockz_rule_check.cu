// blockz_rule_check.cu -- TEST ONLY.  Brute-force check of the depth-plane bound of the RAST_BLOCK_Z variant (kernels.cuh,
// stage_item / raster_item): for random triangles (raster x, y and NDC depth per vertex) and random chunk rectangles, the
// item is staged with the kernel's own stage_item, and for EVERY pixel of the rectangle that the exact path accepts
// (edges / candidate / fragment, i.e. the reference's barycentric test and depth) the claim
//        z_exact(p)  >=  Zo + gx (p.x - rx0) + gy (p.y - ry0) - M
// is checked with the staged Zo, gx, gy, M, evaluated in fp32 as the kernel does at a block corner.  The largest fraction of
// the margin that any pixel used up is reported (the derivation in stage_item predicts < 1 with room to spare).
// Build: nvcc -std=c++17 -O2 -Xcompiler -ffp-contract=off -DRAST_BLOCK_Z=1 -o blockz_rule_check blockz_rule_check.cu
// Usage: blockz_rule_check <seed> <items>  -> one JSON line, exit 1 on a violation
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#ifndef RAST_BLOCK_Z
#define RAST_BLOCK_Z 1
#endif
#include "../rasteriser_b200/csrc/kernels.cuh"

static uint64_t rng_state;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }
static double uni() { return (double)(rnd() >> 11) * (1.0 / 9007199254740992.0); }
static double range(double a, double b) { return a + (b - a) * uni(); }

int main(int argc, char **argv) {
    const uint64_t seed = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1u;
    const uint64_t n = argc > 2 ? strtoull(argv[2], nullptr, 10) : 100000u;
    rng_state = 0x9E3779B97F4A7C15ull ^ (seed * 0xD1B54A32D192ED03ull + 1u);
    for (int i = 0; i < 8; ++i) rnd();
    using namespace rk;
    StagedTris *stg = new StagedTris();
    uint64_t items = 0, usable_items = 0, accepted = 0, violations = 0;
    double worst = 0.0; // max over accepted pixels of (plane - z) / M
    for (uint64_t it = 0; it < n; ++it) {
        const int regime = (int)(rnd() % 8);
        const double cx = range(0.0, 4000.0), cy = range(0.0, 2000.0);
        float4 v[3];
        double size = 100.0;
        switch (regime) {
        case 0: size = range(20.0, 200.0); break;                 // the overdraw workload's size class
        case 1: size = range(2.0, 40.0); break;                   // mid-size
        case 2: size = range(200.0, 3000.0); break;               // huge: the rectangle is a small part
        default: size = std::pow(10.0, range(0.3, 2.8)); break;
        }
        for (int k = 0; k < 3; ++k) { v[k].x = (float)(cx + size * range(-1, 1)); v[k].y = (float)(cy + size * range(-1, 1)); }
        if (regime == 3) { // sliver: third vertex next to the line through the other two
            const double al = uni(), off = std::pow(10.0, range(-5.0, 0.5));
            v[2].x = (float)(v[0].x + al * ((double)v[1].x - v[0].x) + off * range(-1, 1));
            v[2].y = (float)(v[0].y + al * ((double)v[1].y - v[0].y) + off * range(-1, 1));
        }
        // depths: NDC-like, steep, nearly flat, negative, large, tiny
        const int zr = (int)(rnd() % 6);
        for (int k = 0; k < 3; ++k) {
            double z;
            switch (zr) {
            case 0: z = range(0.85, 0.999); break;
            case 1: z = 0.93 + 1e-5 * range(-1, 1); break;
            case 2: z = range(-3.0, 0.99); break;
            case 3: z = range(-1, 1) * 1e4; break;
            case 4: z = range(-1, 1) * 1e-20; break;
            default: z = range(0.0, 1.0); break;
            }
            v[k].z = (float)z;
            v[k].w = 1.f;
        }
        // a chunk rectangle (<= 32 x 32) somewhere in or next to the triangle's bbox, as k_raster_chunks forms them
        const float minx = fminf(fminf(v[0].x, v[1].x), v[2].x), maxx = fmaxf(fmaxf(v[0].x, v[1].x), v[2].x);
        const float miny = fminf(fminf(v[0].y, v[1].y), v[2].y), maxy = fmaxf(fmaxf(v[0].y, v[1].y), v[2].y);
        const double px0 = range(minx - 8.0, maxx + 8.0), py0 = range(miny - 8.0, maxy + 8.0);
        const uint32_t rx0 = (uint32_t)fmax(0.0, px0), ry0 = (uint32_t)fmax(0.0, py0);
        const uint32_t rx1 = rx0 + (uint32_t)(rnd() % 32), ry1 = ry0 + (uint32_t)(rnd() % 32);
        stage_item(*stg, 0u, 7u, 0u, v[0], v[1], v[2], rx0, ry0, rx1, ry1, rx0, ry0);
        ++items;
        const float Zo = exact::u2f(stg->w[27][0]), gx = exact::u2f(stg->w[28][0]), gy = exact::u2f(stg->w[29][0]), M = exact::u2f(stg->w[30][0]);
        if (!(M < 3.0e38f)) continue; // never rejected by the block test
        ++usable_items;
        TriSetup s;
        tri_setup(s, v[0], v[1], v[2]);
        for (uint32_t y = ry0; y <= ry1; ++y)
            for (uint32_t x = rx0; x <= rx1; ++x) {
                float e0, e1, e2, b0, b1, b2, z;
                edges(s, (float)x, (float)y, e0, e1, e2);
                if (!candidate(s, e0, e1, e2)) continue;
                // fragment() also drops z >= 1 (it could never win): the bound must hold for every INSIDE pixel, so redo its first half
                exact::div3(e0, e1, e2, s.area, s.rcp1, s.div_ok, b0, b1, b2);
                if (!(b0 >= 0.f && b1 >= 0.f && b2 >= 0.f)) continue;
                z = exact::add(exact::add(exact::mul(s.z0, b0), exact::mul(s.z1, b1)), exact::mul(s.z2, b2));
                ++accepted;
                const float plane = Zo + gx * (float)(x - rx0) + gy * (float)(y - ry0);
                const double used = ((double)plane - (double)z) / (double)M;
                if (used > worst) worst = used;
                // the shipped test itself (block_behind, used by k_raster_tiles for whole 16 x 8 blocks): a block that holds this pixel must not be
                // declared behind a farthest stored depth equal to the pixel's own depth -- the pixel could still tie or win there
                const uint32_t blk = ((y - ry0) / 8u) * 2u + ((x - rx0) / 16u);
                if (block_behind(blk, depth_key(z), rx0, ry0, rx1, ry1, rx0, ry0, Zo, gx, gy, M)) {
                    if (violations++ < 5) fprintf(stderr, "BLOCK VIOLATION px (%u,%u) block %u z %a\n", x, y, blk, z);
                }
                if (!(z >= plane - M)) {
                    if (violations++ < 5)
                        fprintf(stderr, "VIOLATION px (%u,%u) z %a plane %a M %a v (%a,%a,%a) (%a,%a,%a) (%a,%a,%a)\n", x, y, z, plane, M, v[0].x, v[0].y, v[0].z,
                                v[1].x, v[1].y, v[1].z, v[2].x, v[2].y, v[2].z);
                }
            }
    }
    printf("{\"seed\": %llu, \"items\": %llu, \"usable_items\": %llu, \"accepted_pixels\": %llu, \"violations\": %llu, \"worst_margin_fraction\": %.6g}\n",
           (unsigned long long)seed, (unsigned long long)items, (unsigned long long)usable_items, (unsigned long long)accepted, (unsigned long long)violations, worst);
    delete stg;
    return violations ? 1 : 0;
}
