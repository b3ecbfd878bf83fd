// Kernel launch wrappers of the rain-rendering hot path (sm_100a).  See DESIGN.md for the
// data layout; every wrapper launches on the given stream and returns the CUDA error state.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>
#include "rr_streak_geom.h"

struct rr_frame_bufs {
    // inputs (device)
    const uint8_t *bgr;        // [F][rs*H][rs*W][3]
    const double *bgf;         // [F][3][H][W] reduced float64 image when render_scale == 2, else NULL
    double *bg_sum;            // [F][4] per-channel sum of the (reduced) image in [0,1]
    double *acs;               // [F][4] beta_hg * mean irradiance per channel (k_fog_acs)
    const void *depth;         // [F][H][W] float32 metres, or uint16 PNG samples (metres * 256) when depth_u16
    int depth_u16;
    const rr_streak_rec *streaks;
    const int32_t *offsets;    // [F+1] device copy
    // per-frame intermediates
    unsigned long long *chan_sum;  // [F][4]  sum of uint8 per channel (B,G,R), [3] unused
    double *rainy;             // [F][3][H][W] planar BGR float64
    uint8_t *bg8;              // [F][H][W][4] floor(rainy*255) as (B, G, R, 0): one 32-bit word per pixel
    float *fext;               // extinction exp(-beta d) of the WHOLE batch buffer (not shifted by sub-batch views: frame0 is):
                               // TMA form [F][H + 24][Wp] reflect-padded planes (k_fext_pad), else [F][H][W] (k_fext)
    int frame0;                // first frame of this (sub-)batch inside fext
    int fext_Wp, fext_Hp;      // padded plane size (pitch a multiple of 4 floats = 16 bytes)
    int fog_roll;              // 1: linear frames go through k_fog_roll (strip-walking form), 0: every frame through k_fog
    const float *fext_lut;     // [65536] exp table of the uint16 depth samples (per camera)
    float *fblur;              // [F][H][W] blurred extinction (debug / stage test)
    uint8_t *env8;             // [F][H][env_pitch][4] final environment map, (B, G, R, 0) words; env_pitch = W_env rounded up to 4 pixels
    int env_pitch;             // pixels per row of env8 (16-byte aligned rows: k_env_prefix fetches them with bulk copies)
    int env_bulk;              // k_env_prefix's row fetch: 1 = cp.async.bulk into two buffers, 0 = register-staged
    double *pref;              // [F][H][W_env+1][RR_PREF_N] row prefix sums of (omega*x, omega*y, omega*Y[, omega]), interleaved
    double *rowtot;            // [F][H] row totals of omega*Y
    double *ambient;           // [F] sum over the map of omega*Y
    rr_plan *plans;            // [n_streaks]
    rr_fcp *fcp;               // [n_streaks] prepared field-of-view polygon walkers (k_plan -> k_setup)
    int4 *boxes;               // [n_streaks] (bx0, by0, bw, bh) of the composited block, zeros when nothing is drawn (k_scan)
    int4 *sizes;               // [n_streaks] arena elements (g, v, a) of each streak, written by k_setup for k_scan
    long long *scan;           // [n_streaks+1][6] arena element offsets of g, v, a per streak (3 slots spare); [n][0] = total
    double *arena;             // patch arena
    long long arena_cap;       // elements
    int *err_flag;             // device error flag (arena overflow)
    double *tile_sum;          // [F][rr_n_partials] partial sums of the composited image, one per compositor strip
    double *tile_min, *tile_max;   // [F][rr_n_partials] extrema of the rain mask per compositor strip
    double *frame_mean;        // [F] mean(rainy_bg) - mean(bg)
    double *maskd;             // [F][H][W] float64 rain mask (generator.py:393, bad_weather.py:450)
    double *mask_range;        // [F][2] (min, max) of the float64 rain mask: what plt.imsave normalises with (generator.py:467)
    // outputs (device)
    float *out_bgr;            // [F][H][W][3]
    float *out_mask;           // [F][H][W]
    uint8_t *out_u8;           // [F][H][W][3]
    uint8_t *out_idx8;         // [F][H][W] colormap index of plt.imsave(mask): min(int((m - lo) / (hi - lo) * 256), 255)
    uint16_t *out_u16;         // [F][H][W] the mask min/max-normalised to 16 bits: int((m - lo) / (hi - lo) * 65535 + 0.5)
};

struct rr_static_tabs {
    const int32_t *env_src;    // [H][W_env]
    const uint8_t *env_written;// [H][W_env]
    const uint8_t *env_tile_hole; // [tiles_y][tiles_x] of k_env_map: 1 when the 64x16 tile holds a never-written pixel
    const double *omega;       // [H][W_env]
    const double *omega_pref;  // [H][W_env+1]
    const double *omega_total; // [1] numpy-order total
    const uint8_t *db;         // concatenated textures
    const uint8_t *dbp;        // the same with a one-texel zero border around each (k_raster's bilinear sampler), built by k_build_padded
    const int32_t *tex_poff;   // [n_tex] byte offset of each bordered texture in dbp
    const int32_t *tex_off;    // [n_tex]
    const int32_t *tex_h;      // [n_tex]
};

#ifndef RR_PREF_N
#define RR_PREF_N 4
#endif

struct rr_fog_consts {
    float neg_beta32;          // float32(-beta_ext)
    double irr_scale_num;      // 4 * N^2
    double irr_den;            // exposure_time * gain * pi
    double beta_hg;
};

// compositor: a warp owns a strip of 32 x RR_COMP_PY pixels, RR_COMP_WARPS strips (stacked) per CTA
#ifndef RR_COMP_PY
#define RR_COMP_PY 2
#endif
#define RR_COMP_WARPS 8
// per-frame partial sums of the composited image, one per strip (also scratch of k_downscale2: at least 256)
static inline size_t rr_n_partials(int W, int H) {
    size_t strips = (size_t)((H + RR_COMP_PY * RR_COMP_WARPS - 1) / (RR_COMP_PY * RR_COMP_WARPS)) * RR_COMP_WARPS;
    size_t n = (size_t)((W + 31) / 32) * strips;
    return n > 256 ? n : 256;
}

cudaError_t rr_upload_constants();
cudaError_t rr_prepare_device();       // per-device function attributes (dynamic shared memory opt-in)
// init-time tables
cudaError_t rr_launch_env_tables(int W, int H, int focal_px, int cyl_w, int min_x, int W_env, int32_t *env_src,
                                 uint8_t *env_written, int32_t *cyl_first /* [H][cyl_w] scratch */, cudaStream_t st);
cudaError_t rr_launch_env_tile_flags(const uint8_t *env_written, uint8_t *tile_hole, int H, int W_env, cudaStream_t st);
cudaError_t rr_launch_build_padded(const uint8_t *db, const int32_t *tex_off, const int32_t *tex_h, const int32_t *tex_poff, int n_tex, int tw,
                                   int max_h, uint8_t *dbp, cudaStream_t st);
cudaError_t rr_launch_omega(int H_env, int W_env, double *omega, double *omega_pref, double *omega_total, cudaStream_t st);
// per batch
cudaError_t rr_launch_stats(const rr_frame_bufs &b, int F, int W, int H, int render_scale, double *bgf_out, cudaStream_t st);
// fmap: tensor map of the padded extinction planes (TMA form of the tile load), or NULL for the register-staged form
cudaError_t rr_launch_fog(const rr_frame_bufs &b, const rr_fog_consts &fc, int F, int W, int H, const CUtensorMap *fmap, const CUtensorMap *fmap_roll,
                          bool fext_done, cudaStream_t st);
cudaError_t rr_launch_fext_pad(const rr_frame_bufs &b, const rr_fog_consts &fc, int F, int W, int H, cudaStream_t st);
cudaError_t rr_launch_fext_lut(float *lut, float neg_beta32, cudaStream_t st);
cudaError_t rr_launch_env(const rr_frame_bufs &b, const rr_static_tabs &t, int F, int W, int H, int W_env, cudaStream_t st);
// k_plan needs only the streak records: it may run on another stream while the frame stages (fog, environment map) run
cudaError_t rr_launch_plan(const rr_frame_bufs &b, const rr_static_tabs &t, const rr_cam_dev &cam, int n_streaks, cudaStream_t st);
cudaError_t rr_launch_setup(const rr_frame_bufs &b, const rr_static_tabs &t, const rr_cam_dev &cam, int F, int n_streaks,
                            cudaStream_t st);
cudaError_t rr_launch_scan(const rr_frame_bufs &b, int n_streaks, cudaStream_t st);
cudaError_t rr_launch_raster(const rr_frame_bufs &b, const rr_static_tabs &t, const rr_cam_dev &cam, int n_streaks, int n_sm,
                             cudaStream_t st);
cudaError_t rr_launch_blur(const rr_frame_bufs &b, int n_streaks, int n_sm, cudaStream_t st);
cudaError_t rr_launch_composite(const rr_frame_bufs &b, const rr_cam_dev &cam, int F, cudaStream_t st);
cudaError_t rr_launch_epilogue(const rr_frame_bufs &b, int F, int W, int H, cudaStream_t st);
// stage helpers
cudaError_t rr_launch_env_prefix_only(const rr_frame_bufs &b, const rr_static_tabs &t, int F, int H, int W_env, cudaStream_t st);
cudaError_t rr_launch_planar_to_bg8(const double *planar, uint8_t *bg8, int F, int W, int H, cudaStream_t st);
