// Host+device restatements of the third-party numeric primitives the reference's hot path
// calls (OpenCV 4.x imgproc/core, SciPy ndimage) -- written from their documented/observed
// behaviour, validated on the CPU against cv2/scipy by tests/test_cvmath_host.py through the
// test-only host build (tests/hostsim).  All arithmetic is IEEE double without FMA
// contraction (nvcc -fmad=false / g++ -ffp-contract=off) so host and device agree bit for bit
// on everything except libm transcendentals.
//
// Reference call sites (astra-vision/rain-rendering):
//   cv2.getPerspectiveTransform / warpPerspective(INTER_CUBIC)   common/generator.py:129-131
//   imutils.rotate_bound -> cv2.warpAffine(INTER_LINEAR)          common/generator.py:163
//   cv2.resize(INTER_AREA)                                        common/generator.py:169
//   cv2.fillConvexPoly                                            common/bad_weather.py:388
//   scipy.ndimage.gaussian_filter                                 common/bad_weather.py:296
#pragma once
#include <math.h>
#include <stdint.h>
#include "rr_types.h"

// ---------------------------------------------------------------------------------------
// rounding helpers (cvRound = round-half-even, cvFloor, cvCeil)
// ---------------------------------------------------------------------------------------
RR_HD int rr_round(double v) {
#if defined(__CUDA_ARCH__)
    return __double2int_rn(v);
#else
    return (int)lrint(v);
#endif
}
RR_HD int rr_floor(double v) { return (int)floor(v); }
RR_HD int rr_ceil(double v) { return (int)ceil(v); }
RR_HD int rr_floorf(float v) { return (int)floorf(v); }
RR_HD int rr_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
RR_HD double rr_mind(double a, double b) { return a < b ? a : b; }
RR_HD double rr_maxd(double a, double b) { return a > b ? a : b; }

// ---------------------------------------------------------------------------------------
// Correctly rounded float64 division through a correctly rounded reciprocal (Markstein):
// q0 = n * rd, then one residual correction with two fused multiply-adds.  With rd = RN(1/d) and
// a faithful q0 the result equals RN(n / d); it is used only where tests/test_cvmath_host.py
// proves it EXHAUSTIVELY over the whole input domain: the environment-map xyY conversion (all
// 2^24 colours), where it replaces the five IEEE divisions per pixel (each a ~40-instruction
// sequence on the GPU).
// ---------------------------------------------------------------------------------------
RR_HD double rr_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}
RR_HD double rr_rcp(double d) {      // RN(1 / d)
#if defined(__CUDA_ARCH__)
    return __drcp_rn(d);
#else
    return 1.0 / d;
#endif
}
RR_HD double rr_div_rcp(double n, double d, double rd) {
    double q = n * rd;
    double r = rr_fma(-q, d, n);
    return rr_fma(r, rd, q);
}

// uint8 / 255.0 (the reference's image normalisation, common/generator.py:352) without the division;
// equal to the IEEE quotient for all 256 inputs (tests/test_cvmath_host.py).
RR_HD double rr_u8_unit(uint8_t v) { return rr_div_rcp((double)v, 255.0, 1.0 / 255.0); }

// RGB -> xyY of one environment-map pixel (reference common/my_utils.py:55-68, row vector x M, then
// NaN -> 0 for black, common/generator.py:408).  bb, gg, rr are the uint8 values / 255.0.
// rr_env_xyY_div is the literal form (IEEE divisions), rr_env_xyY the division-free form the kernel runs.
RR_HD void rr_env_xyY_div(double bb, double gg, double rr, double *x, double *y, double *Y) {
    double X = ((rr * 0.49000 + gg * 0.17697) + bb * 0.00000) / 0.17697;
    double Yv = ((rr * 0.31000 + gg * 0.81240) + bb * 0.01000) / 0.17697;
    double Z = ((rr * 0.20000 + gg * 0.01063) + bb * 0.99000) / 0.17697;
    double S = (X + Yv) + Z;
    double xv = X / S, yv = Yv / S;
    if (!(xv == xv)) xv = 0;
    if (!(yv == yv)) yv = 0;
    *x = xv; *y = yv; *Y = Yv;
}
RR_HD void rr_env_xyY(double bb, double gg, double rr, double *x, double *y, double *Y) {
    const double D = 0.17697, RD = 1.0 / 0.17697;      // RN(1/D), folded by the compiler
    double X = rr_div_rcp((rr * 0.49000 + gg * 0.17697) + bb * 0.00000, D, RD);
    double Yv = rr_div_rcp((rr * 0.31000 + gg * 0.81240) + bb * 0.01000, D, RD);
    double Z = rr_div_rcp((rr * 0.20000 + gg * 0.01063) + bb * 0.99000, D, RD);
    double S = (X + Yv) + Z;
    double rs = rr_rcp(S);
    double xv = rr_div_rcp(X, S, rs), yv = rr_div_rcp(Yv, S, rs);
    if (!(xv == xv)) xv = 0;                             // black pixel: 0 * inf = NaN, like 0 / 0
    if (!(yv == yv)) yv = 0;
    *x = xv; *y = yv; *Y = Yv;
}

// texture fetch: streak textures are uint8 gray; the reference divides by 255.0
// (common/bad_weather.py:252) before every resampling call.
RR_HD double rr_tex(const uint8_t *tex, int tw, int x, int y) { return rr_u8_unit(tex[y * tw + x]); }

// ---------------------------------------------------------------------------------------
// OpenCV interpolation tables (imgwarp.cpp: interpolateCubic / initInterTab2D), float
// ---------------------------------------------------------------------------------------
// cubic 1-D table: tab[i*4 + k], i = 0..31 (x = i/32), A = -0.75
inline void rr_build_cubic_tab(float *tab /* 32*4 */) {
    const float A = -0.75f;
    const float scale = 1.f / RR_INTER_TAB;
    for (int i = 0; i < RR_INTER_TAB; i++) {
        volatile float x = i * scale;   // volatile: forbid excess precision / contraction on the host
        float *c = tab + i * 4;
        volatile float t0 = ((A * (x + 1) - 5 * A) * (x + 1) + 8 * A) * (x + 1) - 4 * A;
        volatile float t1 = ((A + 2) * x - (A + 3)) * x * x + 1;
        volatile float t2 = ((A + 2) * (1 - x) - (A + 3)) * (1 - x) * (1 - x) + 1;
        c[0] = t0; c[1] = t1; c[2] = t2;
        c[3] = 1.f - c[0] - c[1] - c[2];
    }
}

// ---------------------------------------------------------------------------------------
// cv::getPerspectiveTransform (8x8 LU with partial pivoting) + 3x3 inverse (cv::invert)
// ---------------------------------------------------------------------------------------
RR_HD bool rr_lu_solve8(double A[8][8], double b[8]) {
    const int m = 8;
    for (int i = 0; i < m; i++) {
        int k = i;
        for (int j = i + 1; j < m; j++)
            if (fabs(A[j][i]) > fabs(A[k][i])) k = j;
        if (fabs(A[k][i]) < 2.220446049250313e-16 * 100) return false;   // DBL_EPSILON*100
        if (k != i) {
            for (int j = i; j < m; j++) { double t = A[i][j]; A[i][j] = A[k][j]; A[k][j] = t; }
            double t = b[i]; b[i] = b[k]; b[k] = t;
        }
        double d = -1 / A[i][i];
        for (int j = i + 1; j < m; j++) {
            double alpha = A[j][i] * d;
            for (int kk = i + 1; kk < m; kk++) A[j][kk] += alpha * A[i][kk];
            b[j] += alpha * b[i];
        }
    }
    for (int i = m - 1; i >= 0; i--) {
        double s = b[i];
        for (int k = i + 1; k < m; k++) s -= A[i][k] * b[k];
        b[i] = s / A[i][i];
    }
    return true;
}

// src/dst are float points (np.float32 arrays in the reference, bad_weather.py:317-327)
RR_HD bool rr_get_perspective(const float sx[4], const float sy[4], const float dx[4], const float dy[4], double M[9]) {
    double a[8][8], b[8];
    for (int i = 0; i < 4; i++) {
        a[i][0] = a[i + 4][3] = sx[i];
        a[i][1] = a[i + 4][4] = sy[i];
        a[i][2] = a[i + 4][5] = 1;
        a[i][3] = a[i][4] = a[i][5] = a[i + 4][0] = a[i + 4][1] = a[i + 4][2] = 0;
        a[i][6] = (double)(-sx[i] * dx[i]);      // float products, as in OpenCV
        a[i][7] = (double)(-sy[i] * dx[i]);
        a[i + 4][6] = (double)(-sx[i] * dy[i]);
        a[i + 4][7] = (double)(-sy[i] * dy[i]);
        b[i] = dx[i];
        b[i + 4] = dy[i];
    }
    bool ok = rr_lu_solve8(a, b);
    for (int i = 0; i < 8; i++) M[i] = ok ? b[i] : 0.0;
    M[8] = 1.;
    return ok;
}

RR_HD bool rr_invert3x3(const double S[9], double t[9]) {
#define SD(r, c) S[(r) * 3 + (c)]
    double d = SD(0, 0) * (SD(1, 1) * SD(2, 2) - SD(1, 2) * SD(2, 1)) -
               SD(0, 1) * (SD(1, 0) * SD(2, 2) - SD(1, 2) * SD(2, 0)) +
               SD(0, 2) * (SD(1, 0) * SD(2, 1) - SD(1, 1) * SD(2, 0));
    if (d == 0.) { for (int i = 0; i < 9; i++) t[i] = 0; return false; }
    d = 1. / d;
    t[0] = (SD(1, 1) * SD(2, 2) - SD(1, 2) * SD(2, 1)) * d;
    t[1] = (SD(0, 2) * SD(2, 1) - SD(0, 1) * SD(2, 2)) * d;
    t[2] = (SD(0, 1) * SD(1, 2) - SD(0, 2) * SD(1, 1)) * d;
    t[3] = (SD(1, 2) * SD(2, 0) - SD(1, 0) * SD(2, 2)) * d;
    t[4] = (SD(0, 0) * SD(2, 2) - SD(0, 2) * SD(2, 0)) * d;
    t[5] = (SD(0, 2) * SD(1, 0) - SD(0, 0) * SD(1, 2)) * d;
    t[6] = (SD(1, 0) * SD(2, 1) - SD(1, 1) * SD(2, 0)) * d;
    t[7] = (SD(0, 1) * SD(2, 0) - SD(0, 0) * SD(2, 1)) * d;
    t[8] = (SD(0, 0) * SD(1, 1) - SD(0, 1) * SD(1, 0)) * d;
#undef SD
    return true;
}

// warpPerspective processes the destination in blocks; the x origin of the block enters the
// floating-point evaluation order (imgwarp.cpp WarpPerspectiveInvoker).
RR_HD int rr_warp_persp_bw0(int width, int height) {
    const int BLOCK_SZ = 32;
    int bh0 = BLOCK_SZ / 2 < height ? BLOCK_SZ / 2 : height;
    int bw0 = BLOCK_SZ * BLOCK_SZ / bh0 < width ? BLOCK_SZ * BLOCK_SZ / bh0 : width;
    return bw0;
}

// One destination pixel of cv2.warpPerspective(tex/255.0, M, (w,h), INTER_CUBIC) with
// BORDER_CONSTANT 0, then np.clip(.,0,1) (generator.py:130-132).  Minv = inverse of M.
RR_HD double rr_warp_persp_cubic(const uint8_t *tex, int tw, int th, const double Minv[9], int bw0,
                                 const float *ctab, int dx, int dy) {
    int xb = (dx / bw0) * bw0, x1 = dx - xb;
    double X0 = Minv[0] * xb + Minv[1] * dy + Minv[2];
    double Y0 = Minv[3] * xb + Minv[4] * dy + Minv[5];
    double W0 = Minv[6] * xb + Minv[7] * dy + Minv[8];
    double W = W0 + Minv[6] * x1;
    W = W ? RR_INTER_TAB / W : 0;
    double fX = rr_maxd(-2147483648.0, rr_mind(2147483647.0, (X0 + Minv[0] * x1) * W));
    double fY = rr_maxd(-2147483648.0, rr_mind(2147483647.0, (Y0 + Minv[3] * x1) * W));
    int X = rr_round(fX), Y = rr_round(fY);
    int sx = rr_clampi(X >> RR_INTER_BITS, -32768, 32767) - 1;
    int sy = rr_clampi(Y >> RR_INTER_BITS, -32768, 32767) - 1;
    const float *wx = ctab + (X & (RR_INTER_TAB - 1)) * 4;
    const float *wy = ctab + (Y & (RR_INTER_TAB - 1)) * 4;
    double out;
    unsigned width1 = tw - 3 > 0 ? tw - 3 : 0, height1 = th - 3 > 0 ? th - 3 : 0;
    if ((unsigned)sx < width1 && (unsigned)sy < height1) {
        double sum = 0;
        for (int i = 0; i < 4; i++) {
            const uint8_t *S = tex + (sy + i) * tw + sx;
            float w0 = wy[i] * wx[0], w1 = wy[i] * wx[1], w2 = wy[i] * wx[2], w3 = wy[i] * wx[3];
            double r = rr_u8_unit(S[0]) * w0 + rr_u8_unit(S[1]) * w1 + rr_u8_unit(S[2]) * w2 +
                       rr_u8_unit(S[3]) * w3;
            sum = (i == 0) ? r : sum + r;
        }
        out = sum;
    } else if (sx >= tw || sx + 4 <= 0 || sy >= th || sy + 4 <= 0) {
        out = 0.0;
    } else {
        double sum = 0.0;   // cval * ONE
        for (int i = 0; i < 4; i++) {
            int yi = sy + i;
            if (yi < 0 || yi >= th) continue;
            for (int j = 0; j < 4; j++) {
                int xj = sx + j;
                if (xj < 0 || xj >= tw) continue;
                float w = wy[i] * wx[j];
                sum += (rr_u8_unit(tex[yi * tw + xj]) - 0.0) * w;
            }
        }
        out = sum;
    }
    return out < 0 ? 0 : (out > 1 ? 1 : out);
}

// ---------------------------------------------------------------------------------------
// imutils.rotate_bound -> cv::getRotationMatrix2D + cv::warpAffine(INTER_LINEAR)
// ---------------------------------------------------------------------------------------
// angle_deg = theta + noise (generator.py:163); rotate_bound negates it.
RR_HD void rr_rotate_bound_setup(int tw, int th, double angle_deg, double Minv[6], int *nW, int *nH) {
    double cX = tw / 2.0, cY = th / 2.0;
    double ang = -angle_deg;
    ang *= 3.1415926535897932384626433832795 / 180;   // CV_PI
    double alpha = cos(ang), beta = sin(ang);          // scale 1.0
    double M[6];
    M[0] = alpha; M[1] = beta; M[2] = (1 - alpha) * cX - beta * cY;
    M[3] = -beta; M[4] = alpha; M[5] = beta * cX + (1 - alpha) * cY;
    double c = fabs(M[0]), s = fabs(M[1]);
    *nW = (int)((th * s) + (tw * c));
    *nH = (int)((th * c) + (tw * s));
    M[2] += (*nW / 2.0) - cX;
    M[5] += (*nH / 2.0) - cY;
    // cv::warpAffine inverts the forward map
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1. / D : 0;
    double A11 = M[4] * D, A22 = M[0] * D;
    M[0] = A11; M[1] *= -D; M[3] *= -D; M[4] = A22;
    double b1 = -M[0] * M[2] - M[1] * M[5];
    double b2 = -M[3] * M[2] - M[4] * M[5];
    M[2] = b1; M[5] = b2;
    for (int i = 0; i < 6; i++) Minv[i] = M[i];
}

// One pixel (x, y) of the rotated canvas: fixed-point source coordinates (AB_BITS = 10,
// INTER_BITS = 5), bilinear float weights, BORDER_CONSTANT 0 (remapBilinear).
RR_HD double rr_warp_affine_linear(const uint8_t *tex, int tw, int th, const double M[6], int x, int y) {
    const int AB_SCALE = 1 << 10;
    int adelta = rr_round(M[0] * x * AB_SCALE);
    int bdelta = rr_round(M[3] * x * AB_SCALE);
    int X0 = rr_round((M[1] * y + M[2]) * AB_SCALE) + AB_SCALE / RR_INTER_TAB / 2;
    int Y0 = rr_round((M[4] * y + M[5]) * AB_SCALE) + AB_SCALE / RR_INTER_TAB / 2;
    int X = (X0 + adelta) >> (10 - RR_INTER_BITS);
    int Y = (Y0 + bdelta) >> (10 - RR_INTER_BITS);
    int sx = rr_clampi(X >> RR_INTER_BITS, -32768, 32767);
    int sy = rr_clampi(Y >> RR_INTER_BITS, -32768, 32767);
    int fx = X & (RR_INTER_TAB - 1), fy = Y & (RR_INTER_TAB - 1);
    const float s = 1.f / RR_INTER_TAB;
    float ax1 = fx * s, ax0 = 1.f - ax1, ay1 = fy * s, ay0 = 1.f - ay1;
    float w0 = ay0 * ax0, w1 = ay0 * ax1, w2 = ay1 * ax0, w3 = ay1 * ax1;
    unsigned width1 = tw - 1 > 0 ? tw - 1 : 0, height1 = th - 1 > 0 ? th - 1 : 0;
    if ((unsigned)sx < width1 && (unsigned)sy < height1) {
        const uint8_t *S = tex + sy * tw + sx;
        return rr_u8_unit(S[0]) * w0 + rr_u8_unit(S[1]) * w1 + rr_u8_unit(S[tw]) * w2 +
               rr_u8_unit(S[tw + 1]) * w3;
    }
    if (sx >= tw || sx + 1 < 0 || sy >= th || sy + 1 < 0) return 0.0;
    bool x0ok = sx >= 0 && sx < tw, x1ok = sx + 1 >= 0 && sx + 1 < tw;
    bool y0ok = sy >= 0 && sy < th, y1ok = sy + 1 >= 0 && sy + 1 < th;
    double v0 = (x0ok && y0ok) ? rr_u8_unit(tex[sy * tw + sx]) : 0.0;
    double v1 = (x1ok && y0ok) ? rr_u8_unit(tex[sy * tw + sx + 1]) : 0.0;
    double v2 = (x0ok && y1ok) ? rr_u8_unit(tex[(sy + 1) * tw + sx]) : 0.0;
    double v3 = (x1ok && y1ok) ? rr_u8_unit(tex[(sy + 1) * tw + sx + 1]) : 0.0;
    return v0 * w0 + v1 * w1 + v2 * w2 + v3 * w3;
}

// Conservative column range of one row of the rotated canvas outside of which every bilinear tap falls
// outside the texture (the sample is then exactly the border constant 0).  The fixed-point source column of
// canvas column c is (B + round(a1024 * c)) >> 10 with a1024 = M * 1024 (see rr_warp_affine_linear); it can
// touch the texture only when lo <= B + round(a1024 * c) <= hi with lo = -1024, hi = size * 1024 - 1.
// round() moves the value by at most 0.5 and the quotient (a multiplication by the reciprocal inv = 1 / a1024,
// off by far less than one column) is widened by one column on both sides, so the returned inclusive range
// [*cmin, *cmax] is a superset of the touching columns (checked against the sampling predicate itself by
// tests/test_cvmath_host.py).
RR_HD double rr_canvas_inv(double a1024) { return (a1024 >= 1.0 || a1024 <= -1.0) ? 1.0 / a1024 : 0.0; }
RR_HD void rr_canvas_axis_range(double a1024, double inv, int B, int size, int *cmin, int *cmax) {
    const double lo = -1024.0, hi = (double)size * 1024.0 - 1.0;
    if (a1024 >= 1.0) {
        *cmin = (int)floor((lo - B - 0.5) * inv) - 1;
        *cmax = (int)floor((hi - B + 0.5) * inv) + 1;
    } else if (a1024 <= -1.0) {
        *cmin = (int)floor((hi - B + 0.5) * inv) - 1;
        *cmax = (int)floor((lo - B - 0.5) * inv) + 1;
    } else {
        *cmin = -0x3fffffff; *cmax = 0x3fffffff;     // (nearly) constant along the row: no restriction
    }
}
// -> inclusive column range [*c_lo, *c_hi] of row (XR, YR) that can be non-zero (empty when c_lo > c_hi);
// inv0 / inv3 = rr_canvas_inv(M[0] * 1024) / rr_canvas_inv(M[3] * 1024)
RR_HD void rr_canvas_row_span(const double *M, double inv0, double inv3, int XR, int YR, int nW, int tw, int th, int *c_lo, int *c_hi) {
    int ax, bx, ay, by;
    rr_canvas_axis_range(M[0] * 1024.0, inv0, XR, tw, &ax, &bx);
    rr_canvas_axis_range(M[3] * 1024.0, inv3, YR, th, &ay, &by);
    int a = ax > ay ? ax : ay, b = bx < by ? bx : by;
    if (a < 0) a = 0;
    if (b > nW - 1) b = nW - 1;
    *c_lo = a; *c_hi = b;
}

// ---------------------------------------------------------------------------------------
// cv::resize(INTER_AREA) on float64: mode selection and one destination pixel per mode.
// The source is addressed through a functor  src(sx, sy) -> double  (the rotated canvas is
// never materialised: each source pixel is recomputed by rr_warp_affine_linear).
// ---------------------------------------------------------------------------------------
#define RR_RESIZE_COPY 0
#define RR_RESIZE_AREA_FAST 1
#define RR_RESIZE_AREA 2
#define RR_RESIZE_LINEAR 3

RR_HD int rr_resize_mode(int sw, int sh, int dw, int dh, double *scale_x, double *scale_y) {
    double inv_x = (double)dw / sw, inv_y = (double)dh / sh;
    double sx = 1. / inv_x, sy = 1. / inv_y;
    *scale_x = sx; *scale_y = sy;
    if (sw == dw && sh == dh) return RR_RESIZE_COPY;
    int ix = rr_round(sx), iy = rr_round(sy);
    bool fast = fabs(sx - ix) < 2.220446049250313e-16 && fabs(sy - iy) < 2.220446049250313e-16;
    if (sx >= 1 && sy >= 1) return fast ? RR_RESIZE_AREA_FAST : RR_RESIZE_AREA;
    return RR_RESIZE_LINEAR;
}

template <class Src>
RR_HD double rr_resize_area_fast(const Src &src, int iscale_x, int iscale_y, int dx, int dy) {
    int area = iscale_x * iscale_y;
    float scale = 1.f / area;
    int x0 = dx * iscale_x, y0 = dy * iscale_y;
    double sum = 0;
    int k = 0;
    // ofs[k] enumerates the cell row-major; OpenCV sums it unrolled by four
    for (; k <= area - 4; k += 4) {
        double a = src(x0 + (k % iscale_x), y0 + (k / iscale_x));
        double b = src(x0 + ((k + 1) % iscale_x), y0 + ((k + 1) / iscale_x));
        double c = src(x0 + ((k + 2) % iscale_x), y0 + ((k + 2) / iscale_x));
        double d = src(x0 + ((k + 3) % iscale_x), y0 + ((k + 3) / iscale_x));
        sum += a + b + c + d;
    }
    for (; k < area; k++) sum += src(x0 + (k % iscale_x), y0 + (k / iscale_x));
    return sum * scale;
}

// DecimateAlpha entries of one destination index (computeResizeAreaTab)
struct rr_area_span { int s_first; int n; float a_first, a_mid, a_last; int has_first, has_last; };

RR_HD rr_area_span rr_area_tab(int d, double scale, int ssize) {
    rr_area_span t;
    double fsx1 = d * scale, fsx2 = fsx1 + scale;
    double cellWidth = rr_mind(scale, ssize - fsx1);
    int sx1 = rr_ceil(fsx1), sx2 = rr_floor(fsx2);
    sx2 = sx2 < ssize - 1 ? sx2 : ssize - 1;
    sx1 = sx1 < sx2 ? sx1 : sx2;
    // the partial-pixel weights are divided out only when they exist: a double division with a zero numerator
    // leaves the GPU's inline fast path for a ~70-instruction subroutine
    t.has_first = (sx1 - fsx1 > 1e-3);
    t.a_first = 0.f;
    if (t.has_first) t.a_first = (float)((sx1 - fsx1) / cellWidth);
    t.s_first = sx1;
    t.n = sx2 - sx1 > 0 ? sx2 - sx1 : 0;
    t.a_mid = (float)(1.0 / cellWidth);
    t.has_last = (fsx2 - sx2 > 1e-3);
    t.a_last = 0.f;
    if (t.has_last) t.a_last = (float)(rr_mind(rr_mind(fsx2 - sx2, 1.), cellWidth) / cellWidth);
    return t;
}

template <class Src>
RR_HD double rr_area_row(const Src &src, const rr_area_span &tx, int sy) {
    double buf = 0;
    if (tx.has_first) buf += src(tx.s_first - 1, sy) * tx.a_first;
    for (int i = 0; i < tx.n; i++) buf += src(tx.s_first + i, sy) * tx.a_mid;
    if (tx.has_last) buf += src(tx.s_first + tx.n, sy) * tx.a_last;
    return buf;
}

template <class Src>
RR_HD double rr_resize_area(const Src &src, int sw, int sh, double scale_x, double scale_y, int dx, int dy) {
    rr_area_span tx = rr_area_tab(dx, scale_x, sw);
    rr_area_span ty = rr_area_tab(dy, scale_y, sh);
    double sum = 0;
    bool first = true;
    if (ty.has_first) {
        double v = ty.a_first * rr_area_row(src, tx, ty.s_first - 1);
        sum = v; first = false;
    }
    for (int i = 0; i < ty.n; i++) {
        double v = ty.a_mid * rr_area_row(src, tx, ty.s_first + i);
        sum = first ? v : sum + v; first = false;
    }
    if (ty.has_last) {
        double v = ty.a_last * rr_area_row(src, tx, ty.s_first + ty.n);
        sum = first ? v : sum + v; first = false;
    }
    return sum;
}

// INTER_AREA when enlarging in at least one direction: the 2-tap "area mode" linear variant
template <class Src>
RR_HD double rr_resize_linear_area(const Src &src, int sw, int sh, int dw, int dh, double scale_x, double scale_y,
                                   int dx, int dy) {
    double inv_x = (double)dw / sw, inv_y = (double)dh / sh;
    int sx = rr_floor(dx * scale_x);
    float fx = (float)((dx + 1) - (sx + 1) * inv_x);
    fx = fx <= 0 ? 0.f : fx - rr_floorf(fx);
    bool xedge = false;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx + 1 >= sw) {
        xedge = true;               // dx >= xmax: D = S[sx] * 1
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    }
    int sy = rr_floor(dy * scale_y);
    float fy = (float)((dy + 1) - (sy + 1) * inv_y);
    fy = fy <= 0 ? 0.f : fy - rr_floorf(fy);
    int sy0 = sy < 0 ? 0 : (sy >= sh ? sh - 1 : sy);
    int sy1 = sy + 1 < 0 ? 0 : (sy + 1 >= sh ? sh - 1 : sy + 1);
    float ax0 = 1.f - fx, ax1 = fx, b0 = 1.f - fy, b1 = fy;
    double r0, r1;
    if (xedge) {
        r0 = src(sx, sy0) * 1.0;
        r1 = src(sx, sy1) * 1.0;
    } else {
        r0 = src(sx, sy0) * ax0 + src(sx + 1, sy0) * ax1;
        r1 = src(sx, sy1) * ax0 + src(sx + 1, sy1) * ax1;
    }
    return r0 * b0 + r1 * b1;
}

// ---------------------------------------------------------------------------------------
// SciPy gaussian_filter1d weights: radius int(4 sigma + 0.5), exp(-0.5/sigma^2 x^2) normalised
// by numpy's pairwise sum (n < 128: eight running sums, then the tail).
// ---------------------------------------------------------------------------------------
RR_HD int rr_gauss_radius(double sigma) { return (int)(4.0 * sigma + 0.5); }

RR_HD double rr_np_sum_small(const double *a, int n) {
    if (n < 8) {
        double res = 0.;
        for (int i = 0; i < n; i++) res += a[i];
        return res;
    }
    double r[8];
    for (int j = 0; j < 8; j++) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
        for (int j = 0; j < 8; j++) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; i++) res += a[i];
    return res;
}

// w[0..2r] (symmetric); caller provides room for 2r+1 doubles, r <= RR_MAX_GAUSS_R
#define RR_MAX_GAUSS_R 60
RR_HD void rr_gauss_weights(double sigma, int r, double *w) {
    double sigma2 = sigma * sigma;
    double f = -0.5 / sigma2;
    for (int i = -r; i <= r; i++) w[i + r] = exp(f * (double)(i * i));
    double s = rr_np_sum_small(w, 2 * r + 1);
    for (int i = 0; i <= 2 * r; i++) w[i] = w[i] / s;
}

// ---------------------------------------------------------------------------------------
// cv::fillConvexPoly on integer vertices (shift 0, line_type 8): the union of
//   (a) the Bresenham outline of every edge (cv::Line, left-to-right LineIterator, after
//       cv::clipLine against the image), and
//   (b) one span per scanline from the two-walker edge scan.
// The mask is evaluated row by row: rr_fcp_prepare() runs the sequential walker once and
// stores, per side, the edge segments; rr_fcp_row() returns the merged column intervals of a row.
// ---------------------------------------------------------------------------------------
struct rr_fcp_seg { int y0, y1; int64_t x, dx; };   // active for rows y0 <= y < y1, x(y) = x + (y - y0) * dx

struct rr_fcp {
    int npts;
    int vx[RR_MAX_POLY], vy[RR_MAX_POLY];
    int ymin, ymax;          // rows covered by the span scan (ymax already clamped); ymin > ymax: none
    int nseg[2];
    rr_fcp_seg seg[2][RR_MAX_POLY];
    int y_stop;              // the scan stops before this row when it runs out of edges
    // clipped outline edges (left point first)
    int ne;
    int ex0[RR_MAX_POLY], ey0[RR_MAX_POLY], ex1[RR_MAX_POLY], ey1[RR_MAX_POLY];
    int W, H;
    int pad_[2];             // sizeof is a multiple of 16: the struct travels between kernels as int4 words
};
static_assert(sizeof(rr_fcp) % 16 == 0, "rr_fcp is copied as int4");

// cv::clipLine(Size, pt1, pt2) on int64 coordinates; returns false when fully outside
RR_HD bool rr_clip_line(int W, int H, int64_t &x1, int64_t &y1, int64_t &x2, int64_t &y2) {
    if (W <= 0 || H <= 0) return false;
    int64_t right = W - 1, bottom = H - 1;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        int64_t a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (int64_t)((double)(a - y1) * (x2 - x1) / (y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (int64_t)((double)(a - y2) * (x2 - x1) / (y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (int64_t)((double)(a - x1) * (y2 - y1) / (x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (int64_t)((double)(a - x2) * (y2 - y1) / (x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

RR_HD void rr_fcp_prepare(rr_fcp &f, int W, int H) {
    const int XY_SHIFT = 16;
    int npts = f.npts;
    f.W = W; f.H = H;
    // ---- outline edges ----
    f.ne = 0;
    {
        int px = f.vx[npts - 1], py = f.vy[npts - 1];
        for (int i = 0; i < npts; i++) {
            int64_t x1 = px, y1 = py, x2 = f.vx[i], y2 = f.vy[i];
            if (rr_clip_line(W, H, x1, y1, x2, y2)) {
                int k = f.ne++;
                if (x2 < x1) { int64_t t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }   // leftToRight
                f.ex0[k] = (int)x1; f.ey0[k] = (int)y1; f.ex1[k] = (int)x2; f.ey1[k] = (int)y2;
            }
            px = f.vx[i]; py = f.vy[i];
        }
    }
    // ---- span scan ----
    int imin = 0;
    int64_t xmin = f.vx[0], xmax = f.vx[0], ymin = f.vy[0], ymax = f.vy[0];
    for (int i = 0; i < npts; i++) {
        if (f.vy[i] < ymin) { ymin = f.vy[i]; imin = i; }
        if (f.vy[i] > ymax) ymax = f.vy[i];
        if (f.vx[i] > xmax) xmax = f.vx[i];
        if (f.vx[i] < xmin) xmin = f.vx[i];
    }
    f.nseg[0] = f.nseg[1] = 0;
    f.ymin = 0; f.ymax = -1; f.y_stop = 0x7fffffff;
    if (npts < 3 || xmax < 0 || ymax < 0 || xmin >= W || ymin >= H) return;
    if (ymax > H - 1) ymax = H - 1;
    f.ymin = (int)ymin; f.ymax = (int)ymax;
    struct { int idx, di; int ye; } edge[2];
    edge[0].idx = edge[1].idx = imin;
    edge[0].ye = edge[1].ye = (int)ymin;
    edge[0].di = 1; edge[1].di = npts - 1;
    int edges = npts;
    int y = (int)ymin;
    // event-driven replay of the do/while scan: only rows where an edge ends need the walker
    while (y <= (int)ymax) {
        for (int i = 0; i < 2; i++) {
            if (y >= edge[i].ye) {
                int idx0 = edge[i].idx, di = edge[i].di;
                int idx = idx0 + di;
                if (idx >= npts) idx -= npts;
                for (; edges-- > 0;) {
                    int ty = f.vy[idx];
                    if (ty > y) {
                        int64_t xs = (int64_t)f.vx[idx0] << XY_SHIFT, xe = (int64_t)f.vx[idx] << XY_SHIFT;
                        rr_fcp_seg s;
                        s.y0 = y; s.y1 = ty;
                        s.dx = ((xe - xs) * 2 + ((int64_t)ty - y)) / (2 * ((int64_t)ty - y));
                        s.x = xs;
                        // close the previous segment of this side at row y
                        if (f.nseg[i] > 0 && f.seg[i][f.nseg[i] - 1].y1 > y) f.seg[i][f.nseg[i] - 1].y1 = y;
                        f.seg[i][f.nseg[i]++] = s;
                        edge[i].ye = ty;
                        edge[i].idx = idx;
                        break;
                    }
                    idx0 = idx;
                    idx += di;
                    if (idx >= npts) idx -= npts;
                }
            }
        }
        if (edges < 0) { f.y_stop = y; break; }
        int ynext = edge[0].ye < edge[1].ye ? edge[0].ye : edge[1].ye;
        if (ynext <= y) ynext = y + 1;
        y = ynext;
    }
}

// x position of side i at row y (XY_SHIFT fixed point); false when no segment is active
RR_HD bool rr_fcp_side_x(const rr_fcp &f, int i, int y, int64_t *x) {
    // the walker keeps incrementing the last segment until replaced, so use the latest
    // segment that started at or before y
    int k = -1;
    for (int s = 0; s < f.nseg[i]; s++)
        if (f.seg[i][s].y0 <= y) k = s;
    if (k < 0) return false;
    *x = f.seg[i][k].x + (int64_t)(y - f.seg[i][k].y0) * f.seg[i][k].dx;
    return true;
}

// Bresenham (8-connected LineIterator) columns of edge k on row y: [a, b], false if none.
// Coordinates are below 2^14 (rr_set_camera enforces it), so 2*dx*t fits in 32 bits.
RR_HD bool rr_line_row(int x0, int y0, int x1, int y1, int y, int *a, int *b) {
    int dx = x1 - x0;            // >= 0 (left to right)
    int dy = y1 - y0;
    int sy = dy < 0 ? -1 : 1;
    int ady = dy < 0 ? -dy : dy;
    int t = (y - y0) * sy;
    if (t < 0 || t > ady) return false;
    if (ady > dx) {
        // y major: one pixel per row, minor offset after j steps is floor((2*dx*j + ady - 1) / (2*ady))
        int m = dx == 0 ? 0 : (2 * dx * t + ady - 1) / (2 * ady);
        *a = *b = x0 + m;
        return true;
    }
    // x major: minor offset after j steps is m_j = floor((2*ady*j + dx - 1) / (2*dx))
    if (ady == 0) { *a = x0; *b = x1; return true; }
    // smallest j with m_j >= t:  2*ady*j + dx - 1 >= 2*dx*t
    int num = 2 * dx * t - dx + 1;
    int jmin = num <= 0 ? 0 : (num + 2 * ady - 1) / (2 * ady);
    int num2 = 2 * dx * (t + 1) - dx + 1;
    int jnext = num2 <= 0 ? 0 : (num2 + 2 * ady - 1) / (2 * ady);
    int jmax = jnext - 1;
    if (jmax > dx) jmax = dx;
    if (jmin > jmax) return false;
    *a = x0 + jmin; *b = x0 + jmax;
    return true;
}

// Column intervals of the mask on row y, merged and sorted; returns their count (<= RR_MAX_POLY+1)
RR_HD int rr_fcp_row(const rr_fcp &f, int y, int *lo, int *hi) {
    int n = 0;
    if (y < 0 || y >= f.H) return 0;
    if (y >= f.ymin && y <= f.ymax && y < f.y_stop) {
        int64_t xa, xb;
        if (rr_fcp_side_x(f, 0, y, &xa) && rr_fcp_side_x(f, 1, y, &xb)) {
            if (xa > xb) { int64_t t = xa; xa = xb; xb = t; }
            const int64_t half = 1 << 15;
            int xx1 = (int)((xa + half) >> 16), xx2 = (int)((xb + half) >> 16);
            if (xx2 >= 0 && xx1 < f.W) {
                if (xx1 < 0) xx1 = 0;
                if (xx2 >= f.W) xx2 = f.W - 1;
                if (xx1 <= xx2) { lo[n] = xx1; hi[n] = xx2; n++; }
            }
        }
    }
    for (int k = 0; k < f.ne; k++) {
        int a, b;
        if (rr_line_row(f.ex0[k], f.ey0[k], f.ex1[k], f.ey1[k], y, &a, &b)) { lo[n] = a; hi[n] = b; n++; }
    }
    // insertion sort by lo, then merge touching / overlapping intervals
    for (int i = 1; i < n; i++) {
        int l = lo[i], h = hi[i], j = i - 1;
        while (j >= 0 && lo[j] > l) { lo[j + 1] = lo[j]; hi[j + 1] = hi[j]; j--; }
        lo[j + 1] = l; hi[j + 1] = h;
    }
    int m = 0;
    for (int i = 0; i < n; i++) {
        if (m > 0 && lo[i] <= hi[m - 1] + 1) { if (hi[i] > hi[m - 1]) hi[m - 1] = hi[i]; }
        else { lo[m] = lo[i]; hi[m] = hi[i]; m++; }
    }
    return m;
}
