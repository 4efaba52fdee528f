// anoddpm_b200/csrc/simplex.cu -- OpenSimplex 2-D/3-D noise for sm_100a.
//
// Replaces the reference's numba CPU kernels (reference simplex.py:195-208, 211-309,
// 321-850) and the host permutation builder (simplex.py:166-192).  Values are computed
// in IEEE fp64 with NO fma contraction (this TU is compiled with -fmad=false) in exactly
// the reference's operation order, so fields are bit-identical to the reference before the
// final fp32 cast; the permutation/gradient-index path is integer and bit-exact.
//
// Formulation (not a transliteration): every lattice-vertex displacement of the 3-D
// algorithm has the shape  d = ((d0 - i) - k*(1/3)) - j  with tiny integers (i, j) per
// axis and one k per vertex, so the kernel first *classifies* the point (region + the two
// "extra" vertices, pure compare/select logic) and then runs ONE uniform contribution
// routine over a 6- or 8-entry vertex list.  That keeps the divergent part of a warp to a
// few selects instead of three 100-line branches.
//
// Roofline class: FP64-ALU bound (~146 flop per voxel-octave, 4 B written per voxel).
#include "common.cuh"
#include "anoddpm_b200.h"
#include "diffusion_step.cuh"
#include <math.h>

namespace {

constexpr double kStretch3 = -1.0 / 6;  // simplex.py:157
constexpr double kSquish3 = 1.0 / 3;    // simplex.py:158
constexpr double kStretch2 = -0.211324865405187;
constexpr double kSquish2 = 0.366025403784439;

__constant__ signed char c_grad3[72] = {
    -11, 4, 4, -4, 11, 4, -4, 4, 11, 11, 4, 4, 4, 11, 4, 4, 4, 11,
    -11, -4, 4, -4, -11, 4, -4, -4, 11, 11, -4, 4, 4, -11, 4, 4, -4, 11,
    -11, 4, -4, -4, 11, -4, -4, 4, -11, 11, 4, -4, 4, 11, -4, 4, 4, -11,
    -11, -4, -4, -4, -11, -4, -4, -4, -11, 11, -4, -4, 4, -11, -4, 4, -4, -11};
__constant__ signed char c_grad2[16] = {5, 2, 2, 5, -5, 2, -2, 5, 5, -2, 2, -5, -5, -2, -2, -5};

// Shared-memory tables of one seed: perm[256], pgi3[256] (values < 256 -> bytes) + gradients.
struct SeedTables {
    unsigned char perm[256];
    unsigned char pgi3[256];
    float grad3[72];
};

__device__ __forceinline__ void load_tables(SeedTables& s, const int16_t* __restrict__ t) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        s.perm[i] = (unsigned char)t[i];
        s.pgi3[i] = (unsigned char)t[256 + i];
    }
    for (int i = threadIdx.x; i < 72; i += blockDim.x) s.grad3[i] = (float)c_grad3[i];
    __syncthreads();
}

// One vertex: lattice offset (i+j per axis), displacement ((d0 - i) - k/3) - j.
struct Base3 {
    double dx0, dy0, dz0;
    long long xsb, ysb, zsb;
};

__device__ __forceinline__ void contribution(double& value, const SeedTables& s, const Base3& b, int ix, int iy,
                                              int iz, int jx, int jy, int jz, int k) {
    const double ks = (double)k * kSquish3;  // 0, 1/3, 2*(1/3), 3*(1/3) exactly as the reference's constants
    const double dx = ((b.dx0 - (double)ix) - ks) - (double)jx;
    const double dy = ((b.dy0 - (double)iy) - ks) - (double)jy;
    const double dz = ((b.dz0 - (double)iz) - ks) - (double)jz;
    double attn = 2 - dx * dx - dy * dy - dz * dz;
    if (attn > 0) {
        const int xs = (int)((b.xsb + ix + jx) & 0xFF), ys = (int)((b.ysb + iy + jy) & 0xFF),
                  zs = (int)((b.zsb + iz + jz) & 0xFF);
        const int g = s.pgi3[(s.perm[(s.perm[xs] + ys) & 0xFF] + zs) & 0xFF];
        const double e = (double)s.grad3[g] * dx + (double)s.grad3[g + 1] * dy + (double)s.grad3[g + 2] * dz;
        attn *= attn;
        value += attn * attn * e;  // added only when attn > 0, like the reference (keeps signed zeros identical)
    }
}

// packed fixed-vertex lists: 4 bits per vertex = (x | y<<1 | z<<2), k = popcount
__device__ __forceinline__ double noise3(const SeedTables& s, double x, double y, double z) {
    const double stretch = (x + y + z) * kStretch3;
    const double xs = x + stretch, ys = y + stretch, zs = z + stretch;
    const double fx = floor(xs), fy = floor(ys), fz = floor(zs);
    Base3 b;
    b.xsb = __double2ll_rd(xs);
    b.ysb = __double2ll_rd(ys);
    b.zsb = __double2ll_rd(zs);
    const double squish = (double)(b.xsb + b.ysb + b.zsb) * kSquish3;
    const double xb = fx + squish, yb = fy + squish, zb = fz + squish;
    const double xins = xs - fx, yins = ys - fy, zins = zs - fz;
    const double in_sum = xins + yins + zins;
    b.dx0 = x - xb;
    b.dy0 = y - yb;
    b.dz0 = z - zb;

    // classification -> fixed vertex list (4 bits each, low nibble first) + two extra vertices
    unsigned fixed_list;
    int nfixed;
    int e0x, e0y, e0z, e0jx = 0, e0jy = 0, e0jz = 0, e0k;
    int e1x, e1y, e1z, e1jx = 0, e1jy = 0, e1jz = 0, e1k;

    if (in_sum <= 1) {  // tetrahedron at (0,0,0)
        fixed_list = 0x4210u;
        nfixed = 4;
        int ap = 1, bp = 2;
        double as = xins, bs = yins;
        if (as >= bs && zins > bs) { bs = zins; bp = 4; }
        else if (as < bs && zins > as) { as = zins; ap = 4; }
        const double wins = 1 - in_sum;
        if (wins > as || wins > bs) {
            const int c = (bs > as) ? bp : ap;
            e0k = e1k = 0;
            e0x = (c & 1) ? 1 : -1;  e1x = (c & 1) ? 1 : 0;
            if (c & 2) { e0y = e1y = 1; }
            else { e0y = (c & 1) ? -1 : 0;  e1y = (c & 1) ? 0 : -1; }
            e0z = (c & 4) ? 1 : 0;   e1z = (c & 4) ? 1 : -1;
        } else {
            const int c = ap | bp;
            e0k = 2; e1k = 1;
            e0x = (c & 1) ? 1 : 0;  e1x = (c & 1) ? 1 : -1;
            e0y = (c & 2) ? 1 : 0;  e1y = (c & 2) ? 1 : -1;
            e0z = (c & 4) ? 1 : 0;  e1z = (c & 4) ? 1 : -1;
        }
    } else if (in_sum >= 2) {  // tetrahedron at (1,1,1)
        fixed_list = 0x7653u;
        nfixed = 4;
        int ap = 6, bp = 5;
        double as = xins, bs = yins;
        if (as <= bs && zins < bs) { bs = zins; bp = 3; }
        else if (as > bs && zins < as) { as = zins; ap = 3; }
        const double wins = 3 - in_sum;
        if (wins < as || wins < bs) {
            const int c = (bs < as) ? bp : ap;
            e0k = e1k = 3;
            e0x = (c & 1) ? 2 : 0;  e1x = (c & 1) ? 1 : 0;
            e0y = e1y = (c & 2) ? 1 : 0;
            if (c & 2) { if (c & 1) e1jy = 1; else e0jy = 1; }
            e0z = (c & 4) ? 1 : 0;  e1z = (c & 4) ? 2 : 0;
        } else {
            const int c = ap & bp;
            e0k = 1; e1k = 2;
            e0x = (c & 1) ? 1 : 0;  e1x = (c & 1) ? 2 : 0;
            e0y = (c & 2) ? 1 : 0;  e1y = (c & 2) ? 2 : 0;
            e0z = (c & 4) ? 1 : 0;  e1z = (c & 4) ? 2 : 0;
        }
    } else {  // octahedron
        fixed_list = 0x653421u;
        nfixed = 6;
        double as, bs, score;
        int ap, bp;
        bool afar, bfar;
        const double p1 = xins + yins;
        if (p1 > 1) { as = p1 - 1; ap = 3; afar = true; }
        else { as = 1 - p1; ap = 4; afar = false; }
        const double p2 = xins + zins;
        if (p2 > 1) { bs = p2 - 1; bp = 5; bfar = true; }
        else { bs = 1 - p2; bp = 2; bfar = false; }
        const double p3 = yins + zins;
        if (p3 > 1) {
            score = p3 - 1;
            if (as <= bs && as < score) { ap = 6; afar = true; }
            else if (as > bs && bs < score) { bp = 6; bfar = true; }
        } else {
            score = 1 - p3;
            if (as <= bs && as < score) { ap = 1; afar = false; }
            else if (as > bs && bs < score) { bp = 1; bfar = false; }
        }
        if (afar == bfar) {
            if (afar) {
                e0x = e0y = e0z = 1; e0k = 3;
                const int c = ap & bp;
                e1x = e1y = e1z = 0; e1k = 2;
                if (c & 1) e1x = 2; else if (c & 2) e1y = 2; else e1z = 2;
            } else {
                e0x = e0y = e0z = 0; e0k = 0;
                const int c = ap | bp;
                e1x = e1y = e1z = 1; e1k = 1;
                if (!(c & 1)) e1x = -1; else if (!(c & 2)) e1y = -1; else e1z = -1;
            }
        } else {
            const int c1 = afar ? ap : bp, c2 = afar ? bp : ap;
            e0x = e0y = e0z = 1; e0k = 1;
            if (!(c1 & 1)) e0x = -1; else if (!(c1 & 2)) e0y = -1; else e0z = -1;
            e1x = e1y = e1z = 0; e1k = 2;
            if (c2 & 1) e1jx = 2; else if (c2 & 2) e1jy = 2; else e1jz = 2;
        }
    }

    double value = 0;
#pragma unroll 1
    for (int v = 0; v < nfixed; ++v) {
        const int code = (fixed_list >> (4 * v)) & 0xF;
        const int ix = code & 1, iy = (code >> 1) & 1, iz = (code >> 2) & 1;
        contribution(value, s, b, ix, iy, iz, 0, 0, 0, ix + iy + iz);
    }
    contribution(value, s, b, e0x, e0y, e0z, e0jx, e0jy, e0jz, e0k);
    contribution(value, s, b, e1x, e1y, e1z, e1jx, e1jy, e1jz, e1k);
    return value / 103;
}

__device__ __forceinline__ double grad2_dot(const SeedTables& s, long long xs, long long ys, double dx, double dy) {
    const int g = s.perm[(s.perm[(int)(xs & 0xFF)] + (int)(ys & 0xFF)) & 0xFF] & 0x0E;
    return (double)c_grad2[g] * dx + (double)c_grad2[g + 1] * dy;
}

#define ADP_CONTRIB2(XS, YS, DX, DY)                               \
    do {                                                           \
        double a_ = 2 - (DX) * (DX) - (DY) * (DY);                 \
        if (a_ > 0) {                                              \
            a_ *= a_;                                              \
            value += a_ * a_ * grad2_dot(s, (XS), (YS), (DX), (DY)); \
        }                                                          \
    } while (0)

__device__ double noise2(const SeedTables& s, double x, double y) {
    const double stretch = (x + y) * kStretch2;
    const double xs = x + stretch, ys = y + stretch;
    const double fx = floor(xs), fy = floor(ys);
    long long xsb = __double2ll_rd(xs), ysb = __double2ll_rd(ys);
    const double squish = (double)(xsb + ysb) * kSquish2;
    const double xins = xs - fx, yins = ys - fy;
    const double in_sum = xins + yins;
    double dx0 = x - (fx + squish), dy0 = y - (fy + squish);
    double value = 0;
    const double dx1 = dx0 - 1 - kSquish2, dy1 = dy0 - 0 - kSquish2;
    ADP_CONTRIB2(xsb + 1, ysb, dx1, dy1);
    const double dx2 = dx0 - 0 - kSquish2, dy2 = dy0 - 1 - kSquish2;
    ADP_CONTRIB2(xsb, ysb + 1, dx2, dy2);
    long long ex, ey;
    double dxe, dye;
    if (in_sum <= 1) {
        const double zins = 1 - in_sum;
        if (zins > xins || zins > yins) {
            if (xins > yins) { ex = xsb + 1; ey = ysb - 1; dxe = dx0 - 1; dye = dy0 + 1; }
            else { ex = xsb - 1; ey = ysb + 1; dxe = dx0 + 1; dye = dy0 - 1; }
        } else { ex = xsb + 1; ey = ysb + 1; dxe = dx0 - 1 - 2 * kSquish2; dye = dy0 - 1 - 2 * kSquish2; }
    } else {
        const double zins = 2 - in_sum;
        if (zins < xins || zins < yins) {
            if (xins > yins) { ex = xsb + 2; ey = ysb; dxe = dx0 - 2 - 2 * kSquish2; dye = dy0 + 0 - 2 * kSquish2; }
            else { ex = xsb; ey = ysb + 2; dxe = dx0 + 0 - 2 * kSquish2; dye = dy0 - 2 - 2 * kSquish2; }
        } else { dxe = dx0; dye = dy0; ex = xsb; ey = ysb; }
        xsb += 1; ysb += 1;
        dx0 = dx0 - 1 - 2 * kSquish2;
        dy0 = dy0 - 1 - 2 * kSquish2;
    }
    ADP_CONTRIB2(xsb, ysb, dx0, dy0);
    ADP_CONTRIB2(ex, ey, dxe, dye);
    return value / 47;
}

// ---- kernels -------------------------------------------------------------------------------

// Fused multi-octave field.  slice s of the output (nslices x H x W) uses table set
// `table_of(s)` and z coordinate `z_of(s)`; per pixel: sum_o amp_o * noise3(x/f_o, y/f_o, z/f_o),
// amp_0 = 1, amp *= persistence, f /= 2  (reference simplex.py:37-54, 75-93).
template <typename OutT>
__global__ void __launch_bounds__(256) simplex3_octaves_kernel(
    OutT* __restrict__ out, int H, int W, int nslices, const int16_t* __restrict__ tables, int table_mode,
    int table_mod, const long long* __restrict__ z_int, const double* __restrict__ z_f64, int z_div, int octaves,
    double persistence, double frequency, const long long* __restrict__ table_step, long long tables_per_step, int freq_pow2) {
    __shared__ SeedTables s;
    const int slice = blockIdx.y;
    int tab = 0;
    if (table_mode == 0) tab = slice;
    else if (table_mode == 2) tab = slice % table_mod;
    // device-resident step index: lets a captured CUDA graph walk through pre-uploaded per-step seed tables
    if (table_step) tables += (size_t)(*table_step) * (size_t)tables_per_step * 512;
    load_tables(s, tables + (size_t)tab * 512);
    double zval;
    if (z_int) zval = (double)z_int[slice / z_div];
    else if (z_f64) zval = z_f64[slice / z_div];
    else zval = (double)slice;
    const int npix = H * W;
    OutT* o = out + (size_t)slice * npix;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
        const int yi = p / W, xi = p - yi * W;
        double f = frequency, amp = 1.0, acc = 0.0;
        if (freq_pow2) {
            // frequency = +-2^k: every octave's f is a power of two, so c / f == c * (1 / f) bit for bit (both exact scalings)
            // and the three IEEE divisions per octave (~25 instructions each) become multiplications
            double inv = 1.0 / f;
            for (int oc = 0; oc < octaves; ++oc) {
                const double n = noise3(s, (double)xi * inv, (double)yi * inv, zval * inv);
                acc += amp * n;
                inv *= 2;
                amp *= persistence;
            }
        } else {
            for (int oc = 0; oc < octaves; ++oc) {
                const double n = noise3(s, (double)xi / f, (double)yi / f, zval / f);
                acc += amp * n;
                f /= 2;
                amp *= persistence;
            }
        }
        o[p] = (OutT)acc;
    }
}

// One reverse step with the noise generated in place (the 12 B/px form of sample_p): per pixel the 6-octave
// simplex value of (x, y, t_b) under the seed table of (step, sample), cast to fp32 exactly like simplex3_octaves_kernel<float>
// stores it, then the posterior step of p_sample_kernel -- bit-identical to running the two kernels, without the noise tensor.
// Single-channel images (slice = sample).  grid (pixel chunks, B).
__global__ void __launch_bounds__(256) p_sample_simplex_kernel(
    float* __restrict__ sample, float* __restrict__ pred_x0, const float* __restrict__ x_t, const float* __restrict__ eps,
    int H, int W, const int16_t* __restrict__ tables, const long long* __restrict__ table_step, long long tables_per_step,
    const long long* __restrict__ t, const float* __restrict__ tab, int T, int octaves, double persistence, double frequency,
    int freq_pow2) {
    __shared__ SeedTables s;
    const int b = blockIdx.y;
    if (table_step) tables += (size_t)(*table_step) * (size_t)tables_per_step * 512;
    load_tables(s, tables + (size_t)b * 512);
    long long tb = t[b];
    const double zval = (double)tb;               // the noise field uses t as given, like anoddpm_simplex_noise
    const bool t_bad = tb < 0 || tb >= T;         // out-of-range timestep: poison instead of an out-of-bounds table read
    if (t_bad) tb = 0;
    const float poison = t_bad ? __int_as_float(0x7fc00000) : 0.0f;
    const float ra = tab[tb] + poison, rm1 = tab[T + tb], c1 = tab[2 * T + tb], c2 = tab[3 * T + tb] + poison, sg = tab[4 * T + tb];
    const int npix = H * W;
    const size_t base = (size_t)b * npix;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += gridDim.x * blockDim.x) {
        const int yi = p / W, xi = p - yi * W;
        double f = frequency, amp = 1.0, acc = 0.0;
        if (freq_pow2) {
            double inv = 1.0 / f;
            for (int oc = 0; oc < octaves; ++oc) {
                const double n = noise3(s, (double)xi * inv, (double)yi * inv, zval * inv);
                acc += amp * n;
                inv *= 2;
                amp *= persistence;
            }
        } else {
            for (int oc = 0; oc < octaves; ++oc) {
                const double n = noise3(s, (double)xi / f, (double)yi / f, zval / f);
                acc += amp * n;
                f /= 2;
                amp *= persistence;
            }
        }
        float x0;
        const float smp = p_step(x_t[base + p], eps[base + p], (float)acc, ra, rm1, c1, c2, sg, x0);
        sample[base + p] = smp;
        if (pred_x0) pred_x0[base + p] = x0;
    }
}

// Generic grid: out[z][y][x] = noise3(X[x], Y[y], Z[z])  (reference _noise3a, simplex.py:833-840)
__global__ void __launch_bounds__(256) simplex3_grid_kernel(double* __restrict__ out, const double* __restrict__ X,
                                                            int nx, const double* __restrict__ Y, int ny,
                                                            const double* __restrict__ Z, int nz,
                                                            const int16_t* __restrict__ tables) {
    __shared__ SeedTables s;
    load_tables(s, tables);
    const size_t n = (size_t)nx * ny * nz;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        const int xi = (int)(p % nx);
        const int yi = (int)((p / nx) % ny);
        const int zi = (int)(p / ((size_t)nx * ny));
        out[p] = noise3(s, X[xi], Y[yi], Z[zi]);
    }
}

// 2-D grid, row-major out[i][j] = noise2(X[j], Y[i]) for i<ny, j<nx.
__global__ void __launch_bounds__(256) simplex2_grid_kernel(double* __restrict__ out, const double* __restrict__ X,
                                                            int nx, const double* __restrict__ Y, int ny,
                                                            const int16_t* __restrict__ tables) {
    __shared__ SeedTables s;
    load_tables(s, tables);
    const size_t n = (size_t)nx * ny;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(p % nx), i = (int)(p / nx);
        out[p] = noise2(s, X[j], Y[i]);
    }
}

}  // namespace

// ---- C ABI -----------------------------------------------------------------------------------

extern "C" int anoddpm_simplex_init(int64_t seed, int16_t* tables) {
    ADP_REQUIRE(tables != nullptr, "anoddpm_simplex_init: null tables");
    // reference simplex.py:166-192.  LCG arithmetic is mod 2^64 (ctypes c_int64 wrap);
    // r = (seed + 31) mod (i + 1) is a floor-mod of the UNWRAPPED sum, evaluated here via residues.
    int16_t source[256];
    for (int i = 0; i < 256; ++i) source[i] = (int16_t)i;
    uint64_t s = (uint64_t)seed;
    const uint64_t A = 6364136223846793005ULL, C = 1442695040888963407ULL;
    for (int w = 0; w < 3; ++w) s = s * A + C;
    for (int i = 255; i >= 0; --i) {
        s = s * A + C;
        const int64_t sv = (int64_t)s, m = i + 1;
        int64_t r = ((sv % m) + (31 % m)) % m;
        if (r < 0) r += m;
        tables[i] = source[r];
        tables[256 + i] = (int16_t)((source[r] % 24) * 3);
        source[r] = source[i];
    }
    return ANODDPM_OK;
}

extern "C" int anoddpm_simplex_init_batch(const int64_t* seeds, int n, int16_t* tables) {
    ADP_REQUIRE(seeds && tables && n >= 0, "anoddpm_simplex_init_batch: bad arguments");
    for (int i = 0; i < n; ++i) anoddpm_simplex_init(seeds[i], tables + (size_t)i * 512);
    return ANODDPM_OK;
}

static int launch_octaves(void* out, int out_f64, int nslices, int H, int W, const int16_t* tables, int table_mode,
                          int table_mod, const int64_t* z_int, const double* z_f64, int z_div, int octaves,
                          double persistence, double frequency, void* stream, const int64_t* table_step = nullptr,
                          long long tables_per_step = 0) {
    ADP_REQUIRE(out && tables, "simplex octaves: null pointer");
    ADP_REQUIRE(nslices > 0 && nslices <= 65535 && H > 0 && W > 0, "simplex octaves: bad shape %d x %d x %d", nslices,
                H, W);
    ADP_REQUIRE(octaves >= 0 && frequency != 0.0, "simplex octaves: bad octaves/frequency");
    ADP_REQUIRE(z_div >= 1 && table_mod >= 1, "simplex octaves: bad divisors");
    const long long npix = (long long)H * W;
    int bx = adp_ceil_div(npix, 256);
    if (bx > 4096) bx = 4096;
    dim3 grid(bx, nslices);
    cudaStream_t st = (cudaStream_t)stream;
    // power-of-two frequency (the reference's 64): x / f is then an exact scaling and is computed as x * (1 / f).  The range
    // check keeps every octave's f and 1 / f normal numbers.
    int fexp = 0;
    const double fm = frexp(frequency, &fexp);
    const int pow2 = (fm == 0.5 || fm == -0.5) && fexp > -900 + octaves && fexp < 900;
    if (out_f64)
        simplex3_octaves_kernel<double><<<grid, 256, 0, st>>>((double*)out, H, W, nslices, tables, table_mode,
                                                             table_mod, (const long long*)z_int, z_f64, z_div,
                                                             octaves, persistence, frequency,
                                                             (const long long*)table_step, tables_per_step, pow2);
    else
        simplex3_octaves_kernel<float><<<grid, 256, 0, st>>>((float*)out, H, W, nslices, tables, table_mode, table_mod,
                                                            (const long long*)z_int, z_f64, z_div, octaves,
                                                            persistence, frequency, (const long long*)table_step,
                                                            tables_per_step, pow2);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_simplex_noise(float* out, int B, int C, int H, int W, const int16_t* tables, int shared_seed,
                                     const int64_t* t, int octaves, double persistence, double frequency,
                                     void* stream) {
    ADP_REQUIRE(t != nullptr, "anoddpm_simplex_noise: null t");
    ADP_REQUIRE(B > 0 && C > 0, "anoddpm_simplex_noise: bad batch/channels");
    return launch_octaves(out, 0, B * C, H, W, tables, shared_seed ? 2 : 0, C, t, nullptr, C, octaves, persistence,
                          frequency, stream);
}

extern "C" int anoddpm_simplex_noise_steps(float* out, int B, int C, int H, int W, const int16_t* tables_all,
                                           const int64_t* step, int shared_seed, const int64_t* t, int octaves,
                                           double persistence, double frequency, void* stream) {
    ADP_REQUIRE(t != nullptr && step != nullptr, "anoddpm_simplex_noise_steps: null t / step");
    ADP_REQUIRE(B > 0 && C > 0, "anoddpm_simplex_noise_steps: bad batch/channels");
    return launch_octaves(out, 0, B * C, H, W, tables_all, shared_seed ? 2 : 0, C, t, nullptr, C, octaves, persistence,
                          frequency, stream, step, shared_seed ? C : (long long)B * C);
}

extern "C" int anoddpm_p_sample_simplex(float* sample, float* pred_x0, const float* x_t, const float* eps,
                                        const int16_t* tables_all, const int64_t* step, const int64_t* t, const float* tables,
                                        int T, int B, int H, int W, int octaves, double persistence, double frequency,
                                        void* stream) {
    ADP_REQUIRE(sample && x_t && eps && tables_all && t && tables, "anoddpm_p_sample_simplex: null pointer");
    ADP_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && T > 0, "anoddpm_p_sample_simplex: bad sizes");
    ADP_REQUIRE(octaves >= 0 && frequency != 0.0, "anoddpm_p_sample_simplex: bad octaves/frequency");
    const long long npix = (long long)H * W;
    int bx = adp_ceil_div(npix, 256);
    if (bx > 4096) bx = 4096;
    int fexp = 0;
    const double fm = frexp(frequency, &fexp);
    const int pow2 = (fm == 0.5 || fm == -0.5) && fexp > -900 + octaves && fexp < 900;
    p_sample_simplex_kernel<<<dim3(bx, B), 256, 0, (cudaStream_t)stream>>>(sample, pred_x0, x_t, eps, H, W, tables_all,
                                                                          (const long long*)step, (long long)B,
                                                                          (const long long*)t, tables, T, octaves, persistence,
                                                                          frequency, pow2);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_simplex_octaves(void* out, int out_f64, int D, int H, int W, const int16_t* tables,
                                       const double* z, int octaves, double persistence, double frequency,
                                       void* stream) {
    return launch_octaves(out, out_f64, D, H, W, tables, 1, 1, nullptr, z, 1, octaves, persistence, frequency, stream);
}

extern "C" int anoddpm_simplex_grid3(double* out, const double* X, int nx, const double* Y, int ny, const double* Z,
                                     int nz, const int16_t* tables, void* stream) {
    ADP_REQUIRE(out && X && Y && Z && tables, "anoddpm_simplex_grid3: null pointer");
    ADP_REQUIRE(nx > 0 && ny > 0 && nz > 0, "anoddpm_simplex_grid3: bad shape");
    const long long n = (long long)nx * ny * nz;
    int bx = adp_ceil_div(n, 256);
    if (bx > 148 * 16) bx = 148 * 16;
    simplex3_grid_kernel<<<bx, 256, 0, (cudaStream_t)stream>>>(out, X, nx, Y, ny, Z, nz, tables);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}

extern "C" int anoddpm_simplex_grid2(double* out, const double* X, int nx, const double* Y, int ny,
                                     const int16_t* tables, void* stream) {
    ADP_REQUIRE(out && X && Y && tables, "anoddpm_simplex_grid2: null pointer");
    ADP_REQUIRE(nx > 0 && ny > 0, "anoddpm_simplex_grid2: bad shape");
    const long long n = (long long)nx * ny;
    int bx = adp_ceil_div(n, 256);
    if (bx > 148 * 16) bx = 148 * 16;
    simplex2_grid_kernel<<<bx, 256, 0, (cudaStream_t)stream>>>(out, X, nx, Y, ny, tables);
    ADP_LAUNCH_CHECK();
    return ANODDPM_OK;
}
