// Per-streak geometry of the hot path, host+device:
//   * field-of-view cone -> env-map polygon        FovComputation.compute_fov_plane_points
//                                                   (reference common/bad_weather.py:596-704)
//   * polygon /\ env rectangle (Clipper restated)  common/bad_weather.py:363-373
//   * patch plan: warp set-up, defocus, placement  common/generator.py:119-171,
//                                                   common/bad_weather.py:286-329,415-434
#pragma once
#include "rr_cvmath.h"

#define RR_PI 3.141592653589793

struct rr_cam_dev {
    int W, H, H_env, W_env;
    double focal_m, f_number, focus_plane, pix_size, radius, fov_deg, opacity_att;
    double exposure_blend;   // cam_exposure / 1000.   (bad_weather.py:344)
    int db_width, n_tex;
};

RR_HD double rr_pymod(double a, double b) {   // Python / numpy float % for b > 0
    double r = fmod(a, b);
    if (r != 0 && r < 0) r += b;
    return r;
}

RR_HD void rr_rotation_matrix(const double ax[3], double theta, double R[9]) {
    double c = cos(theta), s = sin(theta);
    // (c*I) + s*(K) + ((1-c) * outer(axis, axis)),  K = cross-product matrix  (bad_weather.py:533-538)
    double K[9] = {0, -ax[2], ax[1], ax[2], 0, -ax[0], -ax[1], ax[0], 0};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double I = (i == j) ? 1.0 : 0.0;
            R[i * 3 + j] = ((c * I) + s * K[i * 3 + j]) + ((1 - c) * (ax[i] * ax[j]));
        }
}

RR_HD void rr_vecmat(const double v[3], const double M[9], double out[3]) {
    for (int j = 0; j < 3; j++) out[j] = (v[0] * M[j] + v[1] * M[3 + j]) + v[2] * M[6 + j];
}

// The cone is evaluated in three steps so that a warp can spread the 20 rays over its lanes
// (k_setup) while the host build runs them in sequence -- same arithmetic either way.
struct rr_fov_ctx { double P[3], n[3], v[3]; int ok; };
#define RR_FOV_N 20

RR_HD void rr_fov_begin(const rr_streak_rec &s, double fov_deg, rr_fov_ctx &c) {
    double *P = c.P, *n = c.n;
    P[0] = (s.wp1[0] + s.wp2[0]) / 2;
    P[2] = (s.wp1[1] + s.wp2[1]) / 2;    // y <-> z swap (bad_weather.py:599)
    P[1] = (s.wp1[2] + s.wp2[2]) / 2;
    double nrm = sqrt((P[0] * P[0] + P[1] * P[1]) + P[2] * P[2]);
    n[0] = P[0] / nrm; n[1] = P[1] / nrm; n[2] = P[2] / nrm;
    double theta = (fov_deg / 2) * (RR_PI / 180.0);
    double a = n[0], b = n[1], cc = n[2];
    double d = (P[0] * n[0] + P[1] * n[1]) + P[2] * n[2];
    if (b == 0) b = 0.001;
    double qx = P[1];
    double qz = 0;
    double qy = (-a * qx + d - cc * qz) / b;
    double dq[3] = {P[0] - qx, P[1] - qy, P[2] - qz};
    double dn = sqrt((dq[0] * dq[0] + dq[1] * dq[1]) + dq[2] * dq[2]);
    double u[3] = {dq[0] / dn, dq[1] / dn, dq[2] / dn};
    c.ok = (u[0] == u[0]) && (u[1] == u[1]) && (u[2] == u[2]);            // assert ~isnan(u)
    double rv[3] = {u[1] * n[2] - u[2] * n[1], u[2] * n[0] - u[0] * n[2], u[0] * n[1] - u[1] * n[0]};
    double R[9];
    rr_rotation_matrix(rv, -theta, R);
    rr_vecmat(n, R, c.v);
}

// ray k of the cone -> azimuth (image encoding) and env-map pixel coordinates; false on NaN
RR_HD bool rr_fov_ray(const rr_fov_ctx &c, int k, double radius, int rows, int cols, double *az, double *px, double *py) {
    const double two_pi = 2 * RR_PI;
    const double step = two_pi / RR_FOV_N;
    const double *P = c.P;
    double ang = 0 + k * step;
    double M[9], dv[3];
    rr_rotation_matrix(c.n, ang, M);
    rr_vecmat(c.v, M, dv);
    double qa = dv[0] * dv[0] + dv[1] * dv[1] + dv[2] * dv[2];
    double qb = 2 * dv[0] * P[0] + 2 * dv[1] * P[1] + 2 * dv[2] * P[2];
    double qc = P[0] * P[0] + P[1] * P[1] + P[2] * P[2] - radius * radius;
    double disc = qb * qb - 4 * qa * qc;
    double t1 = (-qb + sqrt(disc)) / (2 * qa);
    double x = P[0] + t1 * dv[0], y = P[1] + t1 * dv[1], z = P[2] + t1 * dv[2];
    double el = atan2(z, sqrt(x * x + y * y));
    double azv = atan2(y, x);
    if (azv < 0) azv += two_pi;
    if (el < 0) el += two_pi;
    if (azv > two_pi) azv -= two_pi;
    if (el > two_pi) el -= two_pi;
    azv = ((two_pi - azv) - RR_PI / 2);
    azv = rr_pymod(azv, two_pi);
    double uu = azv / two_pi;
    el = el + RR_PI / 2;
    el = rr_pymod(el, two_pi);
    double vv = 1. - el / RR_PI;
    *az = azv;
    *px = uu * cols;
    *py = vv * rows;
    return (*px == *px) && (*py == *py);
}

// wrap detection and corner splice (bad_weather.py:667-695); az has RR_FOV_N entries; -> 20, 24 or 0
RR_HD int rr_fov_finish(const double *azin, double *px, double *py, int rows, int cols) {
    const int N = RR_FOV_N;
    int count_true = 0, count_false = 0, pos_true = -1, pos_false = -1;
    for (int k = 0; k < N; k++) {
        double df = azin[(k + 1) % N] - azin[k];
        bool cond = (fabs(df) <= 1e-8) || (df < 0);      // np.isclose(diff, 0) | (diff < 0)
        if (cond) { count_true++; if (pos_true < 0) pos_true = k; }
        else { count_false++; if (pos_false < 0) pos_false = k; }
    }
    if (pos_true < 0 || pos_false < 0) return 0;          // np.where(...)[0][0] raises IndexError
    int pos = -1;
    double c0x = 0, c0y = 0, c1x = 0, c1y = 0, c2x = 0, c2y = 0, c3x = 0, c3y = 0;
    if (count_true == 1) {                                 // "top" splice (bad_weather.py:678-684)
        pos = pos_true;
        c0x = cols; c0y = py[pos];
        c1x = cols; c1y = 0;
        c2x = 0;    c2y = 0;
        c3x = 0;    c3y = py[(pos + 1) % N];
    } else if (count_false == 1) {                         // "bottom" splice (:686-692)
        pos = pos_false;
        c0x = 0;    c0y = py[pos];
        c1x = 0;    c1y = rows;
        c2x = cols; c2y = rows;
        c3x = cols; c3y = py[(pos + 1) % N];
    }
    if (pos < 0) return N;
    for (int k = N - 1; k > pos; k--) { px[k + 4] = px[k]; py[k + 4] = py[k]; }
    px[pos + 1] = c0x; py[pos + 1] = c0y;
    px[pos + 2] = c1x; py[pos + 2] = c1y;
    px[pos + 3] = c2x; py[pos + 3] = c2y;
    px[pos + 4] = c3x; py[pos + 4] = c3y;
    return N + 4;
}

// -> number of polygon vertices (20 or 24), 0 when the reference would raise / produce NaNs
RR_HD int rr_fov_polygon(const rr_streak_rec &s, double radius, double fov_deg, int rows, int cols,
                         double *px, double *py) {
    rr_fov_ctx c;
    rr_fov_begin(s, fov_deg, c);
    if (!c.ok) return 0;
    double az[RR_FOV_N];
    for (int k = 0; k < RR_FOV_N; k++)
        if (!rr_fov_ray(c, k, radius, rows, cols, &az[k], &px[k], &py[k])) return 0;
    return rr_fov_finish(az, px, py, rows, cols);
}

// ---- Clipper restated for "polygon /\ rectangle (0,0)-(cols,rows)", see oracle/clipper_rect.py ----
RR_HD int64_t rr_round_half_away(double v) { return v < 0 ? (int64_t)(v - 0.5) : (int64_t)(v + 0.5); }

RR_HD int rr_clip_halfplane(int *x, int *y, int n, int axis, int bound, bool keep_less) {
    int ox[RR_MAX_POLY], oy[RR_MAX_POLY];
    int m = 0;
    for (int i = 0; i < n; i++) {
        int ax = x[i], ay = y[i], bx = x[(i + 1) % n], by = y[(i + 1) % n];
        int av = axis == 0 ? ax : ay, bv = axis == 0 ? bx : by;
        bool ina = keep_less ? (av <= bound) : (av >= bound);
        bool inb = keep_less ? (bv <= bound) : (bv >= bound);
        if (ina) { if (m >= RR_MAX_POLY) return -1; ox[m] = ax; oy[m] = ay; m++; }
        if (ina != inb) {
            double t = (double)(bound - av) / (double)(bv - av);
            int o;
            if (axis == 0) { o = (int)rr_round_half_away(ay + t * (by - ay)); if (m >= RR_MAX_POLY) return -1; ox[m] = bound; oy[m] = o; m++; }
            else { o = (int)rr_round_half_away(ax + t * (bx - ax)); if (m >= RR_MAX_POLY) return -1; ox[m] = o; oy[m] = bound; m++; }
        }
    }
    for (int i = 0; i < m; i++) { x[i] = ox[i]; y[i] = oy[i]; }
    return m;
}

// float polygon -> cleaned, positively oriented integer polygon with the first vertex repeated
// (exactly what the reference hands to cv2.fillConvexPoly, bad_weather.py:372-373).
// Returns the vertex count including the closing vertex, 0 if nothing remains (streak skipped).
RR_HD int rr_clip_fov_polygon(const double *px, const double *py, int n, int cols, int rows, int *vx, int *vy) {
    if (n <= 0) return 0;
    bool lt0x = false, gtx = false, lt0y = false, gty = false;
    for (int i = 0; i < n; i++) {
        vx[i] = (int)px[i];    // C truncation toward zero, like the cInt cast in pyclipper
        vy[i] = (int)py[i];
    }
    for (int i = 0; i < n; i++) lt0x |= vx[i] < 0;
    if (lt0x) { n = rr_clip_halfplane(vx, vy, n, 0, 0, false); if (n < 0) return 0; }
    for (int i = 0; i < n; i++) gtx |= vx[i] > cols;
    if (n && gtx) { n = rr_clip_halfplane(vx, vy, n, 0, cols, true); if (n < 0) return 0; }
    for (int i = 0; i < n; i++) lt0y |= vy[i] < 0;
    if (n && lt0y) { n = rr_clip_halfplane(vx, vy, n, 1, 0, false); if (n < 0) return 0; }
    for (int i = 0; i < n; i++) gty |= vy[i] > rows;
    if (n && gty) { n = rr_clip_halfplane(vx, vy, n, 1, rows, true); if (n < 0) return 0; }
    // clean: drop duplicate / collinear vertices, rescanning from the start after each removal
    bool changed = true;
    while (changed && n >= 3) {
        changed = false;
        for (int i = 0; i < n; i++) {
            int p = i ? i - 1 : n - 1, q = i + 1 < n ? i + 1 : 0;
            bool dup = (vx[i] == vx[p] && vy[i] == vy[p]) || (vx[i] == vx[q] && vy[i] == vy[q]);
            bool col = (int64_t)(vy[i] - vy[p]) * (vx[q] - vx[i]) == (int64_t)(vx[i] - vx[p]) * (vy[q] - vy[i]);
            if (dup || col) {
                for (int k = i; k < n - 1; k++) { vx[k] = vx[k + 1]; vy[k] = vy[k + 1]; }
                n--;
                changed = true;
                break;
            }
        }
    }
    if (n < 3 || n >= RR_MAX_POLY) return 0;      // the closing vertex needs slot n (24 cone vertices + clipping never get there)
    int64_t a = 0;
    for (int i = 0; i < n; i++) {
        int j = i ? i - 1 : n - 1;
        a += (int64_t)(vx[j] + vx[i]) * (vy[j] - vy[i]);
    }
    if (-a < 0) {   // negative Clipper area: reverse
        for (int i = 0; i < n / 2; i++) {
            int t = vx[i]; vx[i] = vx[n - 1 - i]; vx[n - 1 - i] = t;
            t = vy[i]; vy[i] = vy[n - 1 - i]; vy[n - 1 - i] = t;
        }
    }
    vx[n] = vx[0]; vy[n] = vy[0];
    return n + 1;
}

// ---- patch plan ---------------------------------------------------------------------------
RR_HD double rr_circle_of_confusion_px(double o, const rr_cam_dev &cam) {
    double f = cam.focal_m;
    double result = ((o - cam.focus_plane) * (f * f)) / (o * (cam.focus_plane - f) * cam.f_number);
    return result / cam.pix_size;
}

// Fills the geometric part of the plan (everything except photometry and arena offsets).
// tex_h: height of texture s.tex_idx.  Returns false when the reference would raise.
RR_HD bool rr_plan_patch(const rr_streak_rec &s, const rr_cam_dev &cam, int tex_h, rr_plan &p) {
    const int tw = cam.db_width;
    p.type = s.type;
    p.tex_h = tex_h;
    p.flip = 0; p.nW = p.nH = 0; p.bw0 = 1; p.resize_mode = 0; p.scale_x = p.scale_y = 1;
    for (int i = 0; i < 9; i++) p.M[i] = 0;
    if (s.type == RR_BIG) {
        int x0 = s.ip1[0], x1 = s.ip2[0], y0 = s.ip1[1], y1 = s.ip2[1];
        double d0 = floor(s.iw1), d1 = floor(s.iw2);
        int minx = (x0 < x1 ? x0 : x1); if (minx < 0) minx = 0;
        int miny = (y0 < y1 ? y0 : y1); if (miny < 0) miny = 0;
        double maxx = rr_mind(rr_maxd(x0 + d0, x1 + d1), (double)cam.W);
        int maxy = (y0 > y1 ? y0 : y1); if (maxy > cam.H) maxy = cam.H;
        int sw = (int)(maxx - minx), sh = (int)(maxy - miny);
        p.pw = sw > 1 ? sw : 1;
        p.ph = sh > 1 ? sh : 1;
        p.minx = minx; p.miny = miny;
        const double eps = 0.001;
        float sx[4] = {0.f, (float)tw, (float)tw, 0.f};
        float sy[4] = {0.f, 0.f, (float)tex_h, (float)tex_h};
        float dx[4] = {(float)(x0 - minx), (float)(x0 - minx + d0), (float)(x1 - minx + d1 + eps), (float)(x1 - minx + eps)};
        float dy[4] = {(float)(y0 - miny), (float)(y0 - miny), (float)(y1 - miny), (float)(y1 - miny)};
        double M[9];
        rr_get_perspective(sx, sy, dx, dy, M);
        rr_invert3x3(M, p.M);
        p.bw0 = rr_warp_persp_bw0(p.pw, p.ph);
    } else {
        double d0 = s.ip1[0] - s.ip2[0], d1 = s.ip1[1] - s.ip2[1];
        double n1 = sqrt(d0 * d0 + d1 * d1);
        double dir1y = d1 / n1;
        double dir1x = d0 / n1;
        double dotv = dir1x * 0 + dir1y * -1;
        double theta = acos(dotv) * (180.0 / RR_PI);
        rr_rotate_bound_setup(tw, tex_h, theta + s.noise_deg, p.M, &p.nW, &p.nH);
        p.flip = s.ip2m[0] > cam.W / 2;
        int hh = s.ip2m[1] - s.ip1m[1]; if (hh < 0) hh = -hh; if (hh < 2) hh = 2;
        int ww = s.ip2m[0] - s.ip1m[0]; if (ww < 0) ww = -ww; if (ww < s.max_width + 2) ww = s.max_width + 2;
        p.pw = ww; p.ph = hh;
        p.minx = s.ip1m[0]; p.miny = s.ip1m[1];
        if (p.nW <= 0 || p.nH <= 0) return false;
        p.resize_mode = rr_resize_mode(p.nW, p.nH, ww, hh, &p.scale_x, &p.scale_y);
    }
    // defocus (bad_weather.py:286-298)
    double o = fabs(s.wp1[2]);
    double c = fabs(rr_circle_of_confusion_px(o, cam));
    if (!(c == c) || c > 1e6) return false;       // NaN / inf: int(10*c) raises in the reference
    p.sig_y = c; p.sig_x = c / 2;
    p.shift = (int)(10 * c);
    p.ry = c > 1e-15 ? rr_gauss_radius(c) : 0;
    p.rx = (c / 2) > 1e-15 ? rr_gauss_radius(c / 2) : 0;
    if (p.ry > RR_MAX_GAUSS_R || p.rx > RR_MAX_GAUSS_R) return false;
    // placement (bad_weather.py:418-434).  The reference composites the whole zero-padded block
    // (pw + 2 shift) x (ph + 2 shift); the blurred alpha is exactly 0.0 farther than the kernel radius
    // from the patch (every tap there reads padding), and alpha == 0 leaves image and mask bit-for-bit
    // unchanged, so only the block [shift - r, shift + size + r) is blurred, stored and composited.
    int tx = p.minx - p.shift, ty = p.miny - p.shift;                 // image position of padded (0, 0)
    int BW = p.pw + 2 * p.shift, BH = p.ph + 2 * p.shift;
    bool empty = tx > cam.W || ty > cam.H;                            // reference: negative delta -> empty slices
    int Xa = p.shift - p.rx, Xb = p.shift + p.pw + p.rx;
    int Ya = p.shift - p.ry, Yb = p.shift + p.ph + p.ry;
    if (Xa < 0) Xa = 0;
    if (Ya < 0) Ya = 0;
    if (Xb > BW) Xb = BW;
    if (Yb > BH) Yb = BH;
    if (Xa < -tx) Xa = -tx;                                            // clip to the image
    if (Ya < -ty) Ya = -ty;
    if (Xb > cam.W - tx) Xb = cam.W - tx;
    if (Yb > cam.H - ty) Yb = cam.H - ty;
    int visw = Xb - Xa, vish = Yb - Ya;
    if (empty || visw <= 0 || vish <= 0) { visw = 0; vish = 0; Xa = 0; Ya = 0; }
    p.bx0 = tx + Xa; p.by0 = ty + Ya;
    p.cropx = Xa; p.cropy = Ya;
    p.bw = visw; p.bh = vish;
    // blend constants (bad_weather.py:376,425-427)
    double d_avg = (s.iw1 + s.iw2) / 2.;
    double tau_zero = sqrt(1.16 * 1e-3) / 50;
    double length_opacity = cam.opacity_att * d_avg / (s.length + d_avg);
    double tau_one = cam.exposure_blend * length_opacity;
    p.a_scale = tau_one;                 // per pixel: (alpha * tau_one) / exposure
    p.c_scale = tau_one / tau_zero;
    return true;
}

// value of the pre-blur gray patch at (x, y)  (generator.py:126-171)
struct rr_rot_src {
    const uint8_t *tex; int tw, th; const double *M; int flip, nH;
    RR_HD double operator()(int sx, int sy) const {
        int yy = flip ? (nH - 1 - sy) : sy;
        return rr_warp_affine_linear(tex, tw, th, M, sx, yy);
    }
};

RR_HD double rr_patch_pixel(const rr_plan &p, const uint8_t *tex, int tw, const float *ctab, int x, int y) {
    if (p.type == RR_BIG) return rr_warp_persp_cubic(tex, tw, p.tex_h, p.M, p.bw0, ctab, x, y);
    rr_rot_src src = {tex, tw, p.tex_h, p.M, p.flip, p.nH};
    double v;
    switch (p.resize_mode) {
        case RR_RESIZE_COPY: v = src(x, y); break;
        case RR_RESIZE_AREA_FAST: v = rr_resize_area_fast(src, rr_round(p.scale_x), rr_round(p.scale_y), x, y); break;
        case RR_RESIZE_AREA: v = rr_resize_area(src, p.nW, p.nH, p.scale_x, p.scale_y, x, y); break;
        default: v = rr_resize_linear_area(src, p.nW, p.nH, p.pw, p.ph, p.scale_x, p.scale_y, x, y); break;
    }
    return v < 0 ? 0 : (v > 1 ? 1 : v);
}

// tint per unit alpha (bad_weather.py:379,399-412 on a gray pixel of value 1)
RR_HD void rr_tint(double fov_x, double fov_y, double drop_Y, double *kb, double *kg, double *kr) {
    const double m01 = 0.31000, m11 = 0.81240, m21 = 0.01000, factor = 0.17697;
    double Y1 = ((1.0 * m01 + 1.0 * m11) + 1.0 * m21) / factor;
    double Y = Y1 * drop_Y;
    double X = (Y * fov_x) / fov_y;
    double Z = (Y * (1 - fov_x - fov_y)) / fov_y;
    const double M2[9] = {0.41847, -0.15866, -0.082835, -0.091169, 0.25243, 0.015708, 0.0009209, -0.0025498, 0.1786};
    double r = (X * M2[0] + Y * M2[3]) + Z * M2[6];
    double g = (X * M2[1] + Y * M2[4]) + Z * M2[7];
    double b = (X * M2[2] + Y * M2[5]) + Z * M2[8];
    *kb = b; *kg = g; *kr = r;
}
