// Shared plain-C types of the rain-rendering hot path (host C-ABI and device code).
// The public declarations live in include/rain_b200.h; this header only adds the
// device-side plan structs.
#pragma once
#include <stdint.h>
#include "../../include/rain_b200.h"

#if defined(__CUDACC__)
#define RR_HD __host__ __device__ __forceinline__
#else
#define RR_HD inline
#endif

#define RR_BIG 0
#define RR_MEDIUM 1
#define RR_SMALL 2

#define RR_MAX_POLY 32       // 24 FOV vertices + closing vertex, with slack
#define RR_INTER_BITS 5
#define RR_INTER_TAB 32

// Per-(frame, streak) plan produced by the set-up kernel and consumed by the patch
// rasteriser, the defocus blur and the ordered compositor.
struct rr_plan {
    // --- pre-blur gray patch ---------------------------------------------------------
    int32_t valid;          // 0 = streak skipped (degenerate FOV polygon), reference generator.py:185-189
    int32_t type;           // RR_BIG / RR_MEDIUM / RR_SMALL
    int32_t pw, ph;         // patch width / height before the defocus padding
    int32_t minx, miny;     // minC before the defocus shift (generator.py:127,171)
    int32_t tex_off;        // offset (bytes) of the texture in the device DB
    int32_t tex_h;          // texture height (width is db_width)
    // Big: inverse perspective matrix (dst -> src), row major 3x3, and warp block width
    // other: inverse affine matrix in M[0..5]
    double M[9];
    int32_t bw0;            // warpPerspective x-block width (see rr_cvmath.h)
    int32_t nW, nH;         // rotate_bound canvas
    int32_t flip;           // vertical flip of the rotated canvas (generator.py:165)
    int32_t resize_mode;    // 0 copy, 1 area-fast, 2 area, 3 linear(area-mode)
    double scale_x, scale_y;  // resize scales (src/dst) as cv::resize computes them
    // --- defocus ---------------------------------------------------------------------
    double sig_y, sig_x;    // c and c/2  (bad_weather.py:290-296)
    int32_t shift;          // int(10 c)
    int32_t ry, rx;         // SciPy kernel radii int(4 sigma + 0.5)
    // --- placement -------------------------------------------------------------------
    int32_t bx0, by0;       // top-left of the composited block in the image (after clip)
    int32_t cropx, cropy;   // columns/rows of the blurred patch cut away at the left/top
    int32_t bw, bh;         // composited block size (after clipping to the image)
    int64_t g_off;          // byte offset of the texture's zero-bordered copy in the device DB (k_raster's sampler)
    int64_t a_off;          // element offset of the blurred alpha block (bw x bh) in the arena
    // --- photometry ------------------------------------------------------------------
    double kb, kg, kr;      // tint per unit alpha, BGR
    double a_scale;         // tau_one / exposure_time
    double c_scale;         // tau_one / tau_zero
    double fov_x, fov_y, drop_Y;   // diagnostics (stage parity tests)
};
