// TEST-ONLY host build of the host+device math headers (rain_rendering_b200/csrc/rr_cvmath.h,
// rr_streak_geom.h).  It lets the CPU test-suite check the exact arithmetic the CUDA kernels
// execute against cv2 / scipy / the oracle without a GPU.  Never loaded by the product.
#include <string.h>
#include <vector>
#include "../../rain_rendering_b200/csrc/rr_streak_geom.h"

extern "C" {

int hs_fill_convex_poly(const int *vx, const int *vy, int n, int W, int H, uint8_t *mask) {
    static rr_fcp f;
    if (n > RR_MAX_POLY) return -1;
    f.npts = n;
    for (int i = 0; i < n; i++) { f.vx[i] = vx[i]; f.vy[i] = vy[i]; }
    rr_fcp_prepare(f, W, H);
    int lo[RR_MAX_POLY + 2], hi[RR_MAX_POLY + 2];
    int maxiv = 0;
    for (int y = 0; y < H; y++) {
        int m = rr_fcp_row(f, y, lo, hi);
        if (m > maxiv) maxiv = m;
        for (int k = 0; k < m; k++)
            for (int x = lo[k]; x <= hi[k]; x++)
                if (x >= 0 && x < W) mask[y * W + x] = 1;
    }
    return maxiv;
}

// FOV polygon + clip + mask for one streak; returns vertex count (with closing vertex), 0 = skipped
int hs_fov_mask(const rr_streak_rec *s, double radius, double fov_deg, int rows, int cols, uint8_t *mask,
                double *poly_xy /* 24*2 */, int *npoly, int *ivx, int *ivy) {
    double px[32], py[32];
    int n = rr_fov_polygon(*s, radius, fov_deg, rows, cols, px, py);
    *npoly = n;
    for (int i = 0; i < n; i++) { poly_xy[2 * i] = px[i]; poly_xy[2 * i + 1] = py[i]; }
    if (n == 0) return 0;
    static rr_fcp f;
    int m = rr_clip_fov_polygon(px, py, n, cols, rows, f.vx, f.vy);
    if (m == 0) return 0;
    f.npts = m;
    for (int i = 0; i < m; i++) { ivx[i] = f.vx[i]; ivy[i] = f.vy[i]; }
    rr_fcp_prepare(f, cols, rows);
    int lo[RR_MAX_POLY + 2], hi[RR_MAX_POLY + 2];
    for (int y = 0; y < rows; y++) {
        int k = rr_fcp_row(f, y, lo, hi);
        for (int j = 0; j < k; j++)
            for (int x = lo[j]; x <= hi[j]; x++) mask[y * cols + x] = 1;
    }
    return m;
}

// plan + pre-blur patch; out must hold pw*ph doubles (query with out == NULL first)
int hs_patch(const rr_streak_rec *s, const rr_cam_dev *cam, const uint8_t *tex, int tex_h, rr_plan *plan, double *out,
             int out_cap) {
    static float ctab[32 * 4];
    static bool init = false;
    if (!init) { rr_build_cubic_tab(ctab); init = true; }
    memset(plan, 0, sizeof(*plan));
    bool ok = rr_plan_patch(*s, *cam, tex_h, *plan);
    plan->valid = ok;
    if (!ok) return 0;
    if (!out) return plan->pw * plan->ph;
    if (out_cap < plan->pw * plan->ph) return -1;
    for (int y = 0; y < plan->ph; y++)
        for (int x = 0; x < plan->pw; x++) out[y * plan->pw + x] = rr_patch_pixel(*plan, tex, cam->db_width, ctab, x, y);
    return plan->pw * plan->ph;
}

void hs_gauss_weights(double sigma, int *r, double *w) {
    *r = rr_gauss_radius(sigma);
    rr_gauss_weights(sigma, *r, w);
}

void hs_tint(double fx, double fy, double dY, double *out3) { rr_tint(fx, fy, dY, out3, out3 + 1, out3 + 2); }

int hs_sizeof_plan() { return (int)sizeof(rr_plan); }
int hs_sizeof_rec() { return (int)sizeof(rr_streak_rec); }
int hs_sizeof_cam_dev() { return (int)sizeof(rr_cam_dev); }
}

// exhaustive: division-free xyY (the form k_env_prefix runs) against the literal IEEE divisions, all 2^24 colours
extern "C" long hs_env_xyY_mismatches() {
    long bad = 0;
    for (int b = 0; b < 256; b++)
        for (int g = 0; g < 256; g++)
            for (int r = 0; r < 256; r++) {
                double bb = b / 255.0, gg = g / 255.0, rr = r / 255.0, x0, y0, Y0, x1, y1, Y1;
                rr_env_xyY_div(bb, gg, rr, &x0, &y0, &Y0);
                rr_env_xyY(bb, gg, rr, &x1, &y1, &Y1);
                bad += memcmp(&x0, &x1, 8) != 0 || memcmp(&y0, &y1, 8) != 0 || memcmp(&Y0, &Y1, 8) != 0;
            }
    return bad;
}

// Canvas row spans of k_raster (rr_canvas_row_span) against the sampling predicate of every canvas pixel:
// returns the number of pixels outside the span whose bilinear footprint touches the texture (must be 0);
// stats[0] += canvas pixels, stats[1] += pixels inside the spans, stats[2] += pixels that really touch.
extern "C" long hs_canvas_span_violations(const rr_plan *p, int tw, long *stats) {
    const int AB_SCALE = 1 << 10;
    long bad = 0;
    for (int sy = 0; sy < p->nH; sy++) {
        int yy = p->flip ? (p->nH - 1 - sy) : sy;
        int XR = rr_round((p->M[1] * yy + p->M[2]) * AB_SCALE) + AB_SCALE / RR_INTER_TAB / 2;
        int YR = rr_round((p->M[4] * yy + p->M[5]) * AB_SCALE) + AB_SCALE / RR_INTER_TAB / 2;
        int c0, c1;
        rr_canvas_row_span(p->M, rr_canvas_inv(p->M[0] * 1024.0), rr_canvas_inv(p->M[3] * 1024.0), XR, YR, p->nW, tw, p->tex_h, &c0, &c1);
        int cn = c1 >= c0 ? c1 - c0 + 1 : 0;
        stats[0] += p->nW; stats[1] += cn;
        for (int c = 0; c < p->nW; c++) {
            int sx = (XR + rr_round(p->M[0] * c * AB_SCALE)) >> 10, syy = (YR + rr_round(p->M[3] * c * AB_SCALE)) >> 10;
            bool touch = sx >= -1 && sx <= tw - 1 && syy >= -1 && syy <= p->tex_h - 1;
            stats[2] += touch;
            if (touch && (c < c0 || c >= c0 + cn)) bad++;
        }
    }
    return bad;
}

extern "C" long hs_u8_unit_mismatches() {
    long bad = 0;
    for (int v = 0; v < 256; v++) { double a = (double)v / 255.0, b = rr_u8_unit((uint8_t)v); bad += memcmp(&a, &b, 8) != 0; }
    return bad;
}
extern "C" void hs_env_xyY(double bb, double gg, double rr, double *out3) { rr_env_xyY_div(bb, gg, rr, out3, out3 + 1, out3 + 2); }
