/*
 * oracle/simplex_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, IEEE fp64, no FMA contraction) of the reference's
 * OpenSimplex generator.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may load this library; the product path
 * (anoddpm_b200/csrc/simplex.cu) never links or calls it.
 *
 * Parity is PINNED: tests/golden/simplex_*.npz were produced by importing the
 * unmodified reference (numba) in the build container (tests/golden/make_golden.py)
 * and tests/test_oracle_simplex.py checks this file against them bit-for-bit.
 *
 * Reference lines followed (all /root/reference/simplex.py):
 *   oracle_simplex_init      <- _init / overflow          :166-192
 *   grad3_dot                <- _extrapolate3             :202-208, GRADIENTS3 :116-127
 *   grad2_dot                <- _extrapolate2             :195-199, GRADIENTS2 :103-110
 *   oracle_noise2            <- _noise2                   :211-309
 *   oracle_noise3            <- _noise3                   :321-830
 *   oracle_noise3_grid       <- _noise3a                  :833-840
 *   oracle_octaves_fixed_T   <- rand_3d_fixed_T_octaves   :75-93
 *   oracle_octaves_3d        <- rand_3d_octaves           :37-54
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define SQ 0.3333333333333333 /* 1.0/3  (SQUISH_CONSTANT3, :158) */
static const double STRETCH3 = -1.0 / 6; /* :157 */
static const double SQ1 = 1.0 / 3;
static const double SQ2 = 2 * (1.0 / 3);
static const double SQ3 = 3 * (1.0 / 3);
static const double STRETCH2 = -0.211324865405187; /* :155 */
static const double SQU2 = 0.366025403784439;      /* :156 */

static const int8_t G3[72] = {
    -11, 4, 4, -4, 11, 4, -4, 4, 11,
    11, 4, 4, 4, 11, 4, 4, 4, 11,
    -11, -4, 4, -4, -11, 4, -4, -4, 11,
    11, -4, 4, 4, -11, 4, 4, -4, 11,
    -11, 4, -4, -4, 11, -4, -4, 4, -11,
    11, 4, -4, 4, 11, -4, 4, 4, -11,
    -11, -4, -4, -4, -11, -4, -4, -4, -11,
    11, -4, -4, 4, -11, -4, 4, -4, -11};
static const int8_t G2[16] = {5, 2, 2, 5, -5, 2, -2, 5, 5, -2, 2, -5, -5, -2, -2, -5};

/* :166-192.  seed arithmetic is mod 2^64 read back as signed (ctypes c_int64);
 * r = (seed + 31) mod (i+1) is Python floor-mod on the UNWRAPPED sum. */
void oracle_simplex_init(int64_t seed, int64_t *perm, int64_t *pgi3) {
    int64_t source[256];
    uint64_t s = (uint64_t)seed;
    const uint64_t A = 6364136223846793005ULL, C = 1442695040888963407ULL;
    for (int i = 0; i < 256; ++i) source[i] = i;
    s = s * A + C;
    s = s * A + C;
    s = s * A + C;
    for (int i = 255; i >= 0; --i) {
        s = s * A + C;
        int64_t sv = (int64_t)s, m = i + 1;
        int64_t r = ((sv % m) + (31 % m)) % m;
        if (r < 0) r += m;
        perm[i] = source[r];
        pgi3[i] = (perm[i] % 24) * 3;
        source[r] = source[i];
    }
}

static inline double grad3_dot(const int64_t *perm, const int64_t *pgi3, int64_t xs, int64_t ys,
                               int64_t zs, double dx, double dy, double dz) {
    int64_t idx = pgi3[(perm[(perm[xs & 0xFF] + ys) & 0xFF] + zs) & 0xFF];
    return (double)G3[idx] * dx + (double)G3[idx + 1] * dy + (double)G3[idx + 2] * dz;
}

static inline double grad2_dot(const int64_t *perm, int64_t xs, int64_t ys, double dx, double dy) {
    int64_t idx = perm[(perm[xs & 0xFF] + ys) & 0xFF] & 0x0E;
    return (double)G2[idx] * dx + (double)G2[idx + 1] * dy;
}

/* one lattice-vertex contribution: attn = 2 - |d|^2 ; if attn>0: value += attn^4 * grad.d */
#define CONTRIB3(XS, YS, ZS, DX, DY, DZ)                                              \
    do {                                                                              \
        double a_ = 2 - (DX) * (DX) - (DY) * (DY) - (DZ) * (DZ);                      \
        if (a_ > 0) {                                                                 \
            a_ *= a_;                                                                 \
            value += a_ * a_ * grad3_dot(perm, pgi3, (XS), (YS), (ZS), (DX), (DY), (DZ)); \
        }                                                                             \
    } while (0)

double oracle_noise3(double x, double y, double z, const int64_t *perm, const int64_t *pgi3) {
    double stretch = (x + y + z) * STRETCH3;
    double xs = x + stretch, ys = y + stretch, zs = z + stretch;
    int64_t xsb = (int64_t)floor(xs), ysb = (int64_t)floor(ys), zsb = (int64_t)floor(zs);
    double squish = (double)(xsb + ysb + zsb) * SQ1;
    double xb = (double)xsb + squish, yb = (double)ysb + squish, zb = (double)zsb + squish;
    double xins = xs - (double)xsb, yins = ys - (double)ysb, zins = zs - (double)zsb;
    double in_sum = xins + yins + zins;
    double dx0 = x - xb, dy0 = y - yb, dz0 = z - zb;
    double value = 0;
    /* the two region-dependent "extra" lattice vertices */
    int64_t ex0, ey0, ez0, ex1, ey1, ez1;
    double dxe0, dye0, dze0, dxe1, dye1, dze1;

    if (in_sum <= 1) { /* tetrahedron at (0,0,0)  :354-468 */
        int ap = 1, bp = 2;
        double as = xins, bs = yins;
        if (as >= bs && zins > bs) { bs = zins; bp = 4; }
        else if (as < bs && zins > as) { as = zins; ap = 4; }
        double wins = 1 - in_sum;
        if (wins > as || wins > bs) {
            int c = (bs > as) ? bp : ap;
            if (!(c & 1)) { ex0 = xsb - 1; ex1 = xsb; dxe0 = dx0 + 1; dxe1 = dx0; }
            else { ex0 = ex1 = xsb + 1; dxe0 = dxe1 = dx0 - 1; }
            if (!(c & 2)) {
                ey0 = ey1 = ysb; dye0 = dye1 = dy0;
                if (!(c & 1)) { ey1 -= 1; dye1 += 1; }
                else { ey0 -= 1; dye0 += 1; }
            } else { ey0 = ey1 = ysb + 1; dye0 = dye1 = dy0 - 1; }
            if (!(c & 4)) { ez0 = zsb; ez1 = zsb - 1; dze0 = dz0; dze1 = dz0 + 1; }
            else { ez0 = ez1 = zsb + 1; dze0 = dze1 = dz0 - 1; }
        } else {
            int c = ap | bp;
            if (!(c & 1)) { ex0 = xsb; ex1 = xsb - 1; dxe0 = dx0 - SQ2; dxe1 = dx0 + 1 - SQ1; }
            else { ex0 = ex1 = xsb + 1; dxe0 = dx0 - 1 - SQ2; dxe1 = dx0 - 1 - SQ1; }
            if (!(c & 2)) { ey0 = ysb; ey1 = ysb - 1; dye0 = dy0 - SQ2; dye1 = dy0 + 1 - SQ1; }
            else { ey0 = ey1 = ysb + 1; dye0 = dy0 - 1 - SQ2; dye1 = dy0 - 1 - SQ1; }
            if (!(c & 4)) { ez0 = zsb; ez1 = zsb - 1; dze0 = dz0 - SQ2; dze1 = dz0 + 1 - SQ1; }
            else { ez0 = ez1 = zsb + 1; dze0 = dz0 - 1 - SQ2; dze1 = dz0 - 1 - SQ1; }
        }
        CONTRIB3(xsb, ysb, zsb, dx0, dy0, dz0);
        double dx1 = dx0 - 1 - SQ1, dy1 = dy0 - SQ1, dz1 = dz0 - SQ1;
        CONTRIB3(xsb + 1, ysb, zsb, dx1, dy1, dz1);
        double dx2 = dx0 - SQ1, dy2 = dy0 - 1 - SQ1;
        CONTRIB3(xsb, ysb + 1, zsb, dx2, dy2, dz1);
        double dz3 = dz0 - 1 - SQ1;
        CONTRIB3(xsb, ysb, zsb + 1, dx2, dy1, dz3);
    } else if (in_sum >= 2) { /* tetrahedron at (1,1,1)  :469-586 */
        int ap = 6, bp = 5;
        double as = xins, bs = yins;
        if (as <= bs && zins < bs) { bs = zins; bp = 3; }
        else if (as > bs && zins < as) { as = zins; ap = 3; }
        double wins = 3 - in_sum;
        if (wins < as || wins < bs) {
            int c = (bs < as) ? bp : ap;
            if (c & 1) { ex0 = xsb + 2; ex1 = xsb + 1; dxe0 = dx0 - 2 - SQ3; dxe1 = dx0 - 1 - SQ3; }
            else { ex0 = ex1 = xsb; dxe0 = dxe1 = dx0 - SQ3; }
            if (c & 2) {
                ey0 = ey1 = ysb + 1; dye0 = dye1 = dy0 - 1 - SQ3;
                if (c & 1) { ey1 += 1; dye1 -= 1; }
                else { ey0 += 1; dye0 -= 1; }
            } else { ey0 = ey1 = ysb; dye0 = dye1 = dy0 - SQ3; }
            if (c & 4) { ez0 = zsb + 1; ez1 = zsb + 2; dze0 = dz0 - 1 - SQ3; dze1 = dz0 - 2 - SQ3; }
            else { ez0 = ez1 = zsb; dze0 = dze1 = dz0 - SQ3; }
        } else {
            int c = ap & bp;
            if (c & 1) { ex0 = xsb + 1; ex1 = xsb + 2; dxe0 = dx0 - 1 - SQ1; dxe1 = dx0 - 2 - SQ2; }
            else { ex0 = ex1 = xsb; dxe0 = dx0 - SQ1; dxe1 = dx0 - SQ2; }
            if (c & 2) { ey0 = ysb + 1; ey1 = ysb + 2; dye0 = dy0 - 1 - SQ1; dye1 = dy0 - 2 - SQ2; }
            else { ey0 = ey1 = ysb; dye0 = dy0 - SQ1; dye1 = dy0 - SQ2; }
            if (c & 4) { ez0 = zsb + 1; ez1 = zsb + 2; dze0 = dz0 - 1 - SQ1; dze1 = dz0 - 2 - SQ2; }
            else { ez0 = ez1 = zsb; dze0 = dz0 - SQ1; dze1 = dz0 - SQ2; }
        }
        double dx3 = dx0 - 1 - SQ2, dy3 = dy0 - 1 - SQ2, dz3 = dz0 - SQ2;
        CONTRIB3(xsb + 1, ysb + 1, zsb, dx3, dy3, dz3);
        double dy2 = dy0 - SQ2, dz2 = dz0 - 1 - SQ2;
        CONTRIB3(xsb + 1, ysb, zsb + 1, dx3, dy2, dz2);
        double dx1 = dx0 - SQ2;
        CONTRIB3(xsb, ysb + 1, zsb + 1, dx1, dy3, dz2);
        double dx7 = dx0 - 1 - SQ3, dy7 = dy0 - 1 - SQ3, dz7 = dz0 - 1 - SQ3;
        CONTRIB3(xsb + 1, ysb + 1, zsb + 1, dx7, dy7, dz7);
    } else { /* octahedron in between  :587-798 */
        double as, bs, score;
        int ap, bp, afar, bfar;
        double p1 = xins + yins;
        if (p1 > 1) { as = p1 - 1; ap = 3; afar = 1; }
        else { as = 1 - p1; ap = 4; afar = 0; }
        double p2 = xins + zins;
        if (p2 > 1) { bs = p2 - 1; bp = 5; bfar = 1; }
        else { bs = 1 - p2; bp = 2; bfar = 0; }
        double p3 = yins + zins;
        if (p3 > 1) {
            score = p3 - 1;
            if (as <= bs && as < score) { ap = 6; afar = 1; }
            else if (as > bs && bs < score) { bp = 6; bfar = 1; }
        } else {
            score = 1 - p3;
            if (as <= bs && as < score) { ap = 1; afar = 0; }
            else if (as > bs && bs < score) { bp = 1; bfar = 0; }
        }
        if (afar == bfar) {
            if (afar) { /* both on the (1,1,1) side */
                dxe0 = dx0 - 1 - SQ3; dye0 = dy0 - 1 - SQ3; dze0 = dz0 - 1 - SQ3;
                ex0 = xsb + 1; ey0 = ysb + 1; ez0 = zsb + 1;
                int c = ap & bp;
                dxe1 = dx0 - SQ2; dye1 = dy0 - SQ2; dze1 = dz0 - SQ2;
                ex1 = xsb; ey1 = ysb; ez1 = zsb;
                if (c & 1) { dxe1 = dx0 - 2 - SQ2; ex1 = xsb + 2; }
                else if (c & 2) { dye1 = dy0 - 2 - SQ2; ey1 = ysb + 2; }
                else { dze1 = dz0 - 2 - SQ2; ez1 = zsb + 2; }
            } else { /* both on the (0,0,0) side */
                dxe0 = dx0; dye0 = dy0; dze0 = dz0;
                ex0 = xsb; ey0 = ysb; ez0 = zsb;
                int c = ap | bp;
                dxe1 = dx0 - 1 - SQ1; dye1 = dy0 - 1 - SQ1; dze1 = dz0 - 1 - SQ1;
                ex1 = xsb + 1; ey1 = ysb + 1; ez1 = zsb + 1;
                if (!(c & 1)) { dxe1 = dx0 + 1 - SQ1; ex1 = xsb - 1; }
                else if (!(c & 2)) { dye1 = dy0 + 1 - SQ1; ey1 = ysb - 1; }
                else { dze1 = dz0 + 1 - SQ1; ez1 = zsb - 1; }
            }
        } else { /* one closest point on each side */
            int c1 = afar ? ap : bp, c2 = afar ? bp : ap;
            dxe0 = dx0 - 1 - SQ1; dye0 = dy0 - 1 - SQ1; dze0 = dz0 - 1 - SQ1;
            ex0 = xsb + 1; ey0 = ysb + 1; ez0 = zsb + 1;
            if (!(c1 & 1)) { dxe0 = dx0 + 1 - SQ1; ex0 = xsb - 1; }
            else if (!(c1 & 2)) { dye0 = dy0 + 1 - SQ1; ey0 = ysb - 1; }
            else { dze0 = dz0 + 1 - SQ1; ez0 = zsb - 1; }
            dxe1 = dx0 - SQ2; dye1 = dy0 - SQ2; dze1 = dz0 - SQ2;
            ex1 = xsb; ey1 = ysb; ez1 = zsb;
            if (c2 & 1) { dxe1 -= 2; ex1 += 2; }
            else if (c2 & 2) { dye1 -= 2; ey1 += 2; }
            else { dze1 -= 2; ez1 += 2; }
        }
        double dx1 = dx0 - 1 - SQ1, dy1 = dy0 - SQ1, dz1 = dz0 - SQ1;
        CONTRIB3(xsb + 1, ysb, zsb, dx1, dy1, dz1);
        double dx2 = dx0 - SQ1, dy2 = dy0 - 1 - SQ1;
        CONTRIB3(xsb, ysb + 1, zsb, dx2, dy2, dz1);
        double dz3 = dz0 - 1 - SQ1;
        CONTRIB3(xsb, ysb, zsb + 1, dx2, dy1, dz3);
        double dx4 = dx0 - 1 - SQ2, dy4 = dy0 - 1 - SQ2, dz4 = dz0 - SQ2;
        CONTRIB3(xsb + 1, ysb + 1, zsb, dx4, dy4, dz4);
        double dy5 = dy0 - SQ2, dz5 = dz0 - 1 - SQ2;
        CONTRIB3(xsb + 1, ysb, zsb + 1, dx4, dy5, dz5);
        double dx6 = dx0 - SQ2;
        CONTRIB3(xsb, ysb + 1, zsb + 1, dx6, dy4, dz5);
    }
    CONTRIB3(ex0, ey0, ez0, dxe0, dye0, dze0);
    CONTRIB3(ex1, ey1, ez1, dxe1, dye1, dze1);
    return value / 103;
}

#define CONTRIB2(XS, YS, DX, DY)                                         \
    do {                                                                 \
        double a_ = 2 - (DX) * (DX) - (DY) * (DY);                       \
        if (a_ > 0) {                                                    \
            a_ *= a_;                                                    \
            value += a_ * a_ * grad2_dot(perm, (XS), (YS), (DX), (DY));  \
        }                                                                \
    } while (0)

double oracle_noise2(double x, double y, const int64_t *perm) {
    double stretch = (x + y) * STRETCH2;
    double xs = x + stretch, ys = y + stretch;
    int64_t xsb = (int64_t)floor(xs), ysb = (int64_t)floor(ys);
    double squish = (double)(xsb + ysb) * SQU2;
    double xb = (double)xsb + squish, yb = (double)ysb + squish;
    double xins = xs - (double)xsb, yins = ys - (double)ysb;
    double in_sum = xins + yins;
    double dx0 = x - xb, dy0 = y - yb;
    double value = 0;
    double dx1 = dx0 - 1 - SQU2, dy1 = dy0 - 0 - SQU2;
    CONTRIB2(xsb + 1, ysb, dx1, dy1);
    double dx2 = dx0 - 0 - SQU2, dy2 = dy0 - 1 - SQU2;
    CONTRIB2(xsb, ysb + 1, dx2, dy2);
    int64_t ex, ey;
    double dxe, dye;
    if (in_sum <= 1) {
        double zins = 1 - in_sum;
        if (zins > xins || zins > yins) {
            if (xins > yins) { ex = xsb + 1; ey = ysb - 1; dxe = dx0 - 1; dye = dy0 + 1; }
            else { ex = xsb - 1; ey = ysb + 1; dxe = dx0 + 1; dye = dy0 - 1; }
        } else { ex = xsb + 1; ey = ysb + 1; dxe = dx0 - 1 - 2 * SQU2; dye = dy0 - 1 - 2 * SQU2; }
    } else {
        double zins = 2 - in_sum;
        if (zins < xins || zins < yins) {
            if (xins > yins) { ex = xsb + 2; ey = ysb; dxe = dx0 - 2 - 2 * SQU2; dye = dy0 + 0 - 2 * SQU2; }
            else { ex = xsb; ey = ysb + 2; dxe = dx0 + 0 - 2 * SQU2; dye = dy0 - 2 - 2 * SQU2; }
        } else { dxe = dx0; dye = dy0; ex = xsb; ey = ysb; }
        xsb += 1; ysb += 1;
        dx0 = dx0 - 1 - 2 * SQU2;
        dy0 = dy0 - 1 - 2 * SQU2;
    }
    CONTRIB2(xsb, ysb, dx0, dy0);
    CONTRIB2(ex, ey, dxe, dye);
    return value / 47;
}

/* :833-840 -- out[z][y][x] */
void oracle_noise3_grid(const double *X, int nx, const double *Y, int ny, const double *Z, int nz,
                        const int64_t *perm, const int64_t *pgi3, double *out) {
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x)
                out[((size_t)z * ny + y) * nx + x] = oracle_noise3(X[x], Y[y], Z[z], perm, pgi3);
}

/* :37-54 and :75-93.  noise += amplitude * noise3array(x/f, y/f, z/f); f /= 2; amplitude *= persistence.
 * z coordinates are given explicitly (arange(D) for rand_3d_octaves, the timestep(s) for fixed_T). */
void oracle_octaves(int nx, int ny, int y_begin, int y_end, const double *zvals, int nz, int octaves,
                    double persistence, double frequency, const int64_t *perm, const int64_t *pgi3,
                    double *out) {
    /* rows [y_begin, y_end) of every z-slice; the Python side threads over row ranges
     * (this image has no libgomp) -- the per-voxel arithmetic order is unaffected. */
    for (int z = 0; z < nz; ++z)
        memset(out + ((size_t)z * ny + y_begin) * nx, 0, (size_t)(y_end - y_begin) * nx * sizeof(double));
    double amp = 1, f = frequency;
    for (int o = 0; o < octaves; ++o) {
        for (int z = 0; z < nz; ++z)
            for (int y = y_begin; y < y_end; ++y) {
                double zz = zvals[z] / f, yy = (double)y / f;
                double *row = out + ((size_t)z * ny + y) * nx;
                for (int x = 0; x < nx; ++x)
                    row[x] += amp * oracle_noise3((double)x / f, yy, zz, perm, pgi3);
            }
        f /= 2;
        amp *= persistence;
    }
}
