/* rain_b200.h -- C ABI of the B200-native rain-rendering hot path (librain_b200.so).
 *
 * Drop-in boundary for the per-frame path of astra-vision/rain-rendering:
 *   common/generator.py:299-469  (Generator.run per-frame body)
 *   common/add_attenuation.py:26-95, common/bad_weather.py:272-853,
 *   common/solid_angle.py:5-102, common/my_utils.py:55-85.
 * The reference is pure Python, so the binding a maintainer adds is a ctypes stub
 * (shown in INTEGRATION.md; shipped as rain_rendering_b200/_lib.py).
 *
 * Conventions: every function returns 0 on success and a negative rr_status otherwise;
 * rr_last_error() returns a static/thread-local message.  The caller owns every buffer it
 * passes; the library owns device memory behind the opaque context.  One context is bound
 * to one GPU and one CUDA stream; a context is NOT thread-safe, distinct contexts are.
 * There is no CPU fallback: rr_create fails if no CUDA device is usable.
 */
#ifndef RAIN_B200_H
#define RAIN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rr_context rr_context;

enum rr_status {
    RR_OK = 0,
    RR_ERR_CUDA = -1,        /* a CUDA runtime call failed                     */
    RR_ERR_ARG = -2,         /* invalid argument                               */
    RR_ERR_STATE = -3,       /* call order violated (no camera / no streak DB) */
    RR_ERR_CAPACITY = -4,    /* a device arena overflowed                      */
    RR_PNG_UNSUPPORTED = -5, /* a valid PNG the native codec does not decode (palette, interlace, < 8 bits) */
    RR_PNG_SIZE = -6         /* decoded, but not the expected width x height   */
};

/* One imaged streak of one simulator frame, as common/bad_weather.py:46-60 ("Streak") holds it
 * after load_streaks_from_xml (:200-239) and the in-frame filter (common/generator.py:413-420),
 * plus the two host-side NumPy RNG draws of the frame (legacy MT19937 stream seeded per frame,
 * generator.py:318): the texture index (bad_weather.py:250-265) and the wind noise in degrees
 * (generator.py:136).  Geometry is carried in float64 because the reference computes in
 * float64.  128 bytes, little-endian, naturally aligned. */
typedef struct rr_streak_rec {
    double wp1[3], wp2[3];   /* world start / end, metres, z already negated (bad_weather.py:223) */
    double iw1, iw2;         /* image diameters in px (after /render_scale)                       */
    double noise_deg;        /* np.random.normal(0, noise_std) * noise_scale; 0 for Big drops     */
    double ratio;            /* max_width / actual_length (bad_weather.py:233); informational     */
    int32_t ip1[2], ip2[2];  /* rounded image start / end (x, y), y down, BEFORE the wind rotation */
    int32_t ip1m[2], ip2m[2];/* the same AFTER the wind rotation wrote back (generator.py:152-161)  */
    int32_t max_width;       /* int(max(iw1, iw2))                                                */
    int32_t length;          /* ceil(|ip1 - ip2|)                                                 */
    int32_t pid;
    uint8_t type;            /* 0 Big (max_width >= 4), 1 Medium (> 1), 2 Small                   */
    uint8_t tex_idx;         /* np.random.randint draw: index into the streak DB                  */
    uint8_t pad[2];
} rr_streak_rec;

/* Per (sequence, weather) constants: common/generator.py:51-55,232-233,260-267 and
 * common/bad_weather.py:344,469,712. */
typedef struct rr_camera {
    int32_t W, H;            /* render size                                       */
    double focal_m;          /* cam_focal / 1000                                  */
    double f_number;
    double exposure_ms;
    double gain;
    double focus_plane_m;    /* 6   (hard-coded at generator.py:267)              */
    double pix_size_m;       /* 4.65e-6 (hard-coded at bad_weather.py:469)        */
    double radius;           /* 10  (generator.py:267)                            */
    double fov_deg;          /* 165 (generator.py:267)                            */
    double opacity_att;      /* --opacity_attenuation                             */
    double fallrate_mmh;     /* rain intensity of the weather                     */
    int32_t render_scale;    /* settings["render_scale"]: 1, or 2 = the input frames are (2H, 2W) and are
                              * reduced like cv2.resize does for an exact factor 2 (generator.py:354-355) */
    int32_t reserved;
} rr_camera;

/* timing slots of rr_timings() */
enum { RR_T_H2D = 0, RR_T_FOG, RR_T_ENV, RR_T_SETUP, RR_T_RASTER, RR_T_BLUR, RR_T_COMPOSITE, RR_T_EPILOGUE,
       RR_T_D2H, RR_T_TOTAL, RR_T_COUNT };

/* debug read-back selectors of rr_debug_read() (stage parity tests) */
enum { RR_DBG_FOG_F64 = 0,      /* (3,H,W) float64 planar BGR after the fog stage               */
       RR_DBG_ENV_U8 = 1,       /* (H,W_env,3) uint8 BGR environment map                         */
       RR_DBG_OMEGA = 2,        /* (H,W_env) float64 solid angles                                */
       RR_DBG_PLANS = 3,        /* rr_plan[n_streaks of the frame] (see csrc/rr_types.h)         */
       RR_DBG_RAINY_F64 = 4,    /* (3,H,W) float64 planar BGR after compositing, before mean shift */
       RR_DBG_ENV_SRC = 5,      /* (H,W_env) int32 source index table                            */
       RR_DBG_ARENA = 6,        /* the float64 patch arena (offsets in rr_plan)                  */
       RR_DBG_FEXT = 7 };       /* (H,W) float32 blurred extinction                              */

int rr_version(void);
const char *rr_last_error(void);

int rr_create(int device_id, rr_context **out);
int rr_destroy(rr_context *ctx);

/* Streak DB: n_tex gray textures of common width, concatenated row-major, natural-sort order,
 * normalised uint8 exactly as DBManager.load_streak_database leaves them (bad_weather.py:139-146;
 * the reference's BGR copy has three identical channels, one is passed). */
int rr_set_streak_db(rr_context *ctx, int n_tex, const int32_t *heights, int width, const uint8_t *gray_concat);
/* Multi-GPU: non-root ranks allocate, the caller broadcasts into the device buffer (NCCL),
 * see rain_rendering_b200/dist.py. */
int rr_alloc_streak_db(rr_context *ctx, int n_tex, const int32_t *heights, int width);
int rr_streak_db_device_ptr(rr_context *ctx, void **dev_ptr, size_t *bytes);

/* Builds the per-camera tables on the device: cylindrical source-index map, hole mask,
 * solid angles, their row prefix sums (replaces EnvironmentMapGenerator.__init__ + the
 * geometry half of generate_map, and solid_angle.get_solid_angles which the reference
 * recomputes every frame).  max_batch frames are provisioned. */
int rr_set_camera(rr_context *ctx, const rr_camera *cam, int max_batch);
int rr_env_size(rr_context *ctx, int *H_env, int *W_env);

/* The hot path for a batch of n_frames (<= max_batch) independent frames.  HOST buffers:
 *   bgr        n*(rs*H)*(rs*W)*3 uint8   (cv2.imread order, generator.py:352; rs = render_scale)
 *   depth      n*H*W   float32 metres (generator.py:365)
 *   streaks    concatenated records, frame f owns [streak_offsets[f], streak_offsets[f+1])
 *   out_bgr    n*H*W*3 float32 BGR mean-shifted rainy image (generator.py:464), may be NULL
 *   out_mask   n*H*W   float32 rain mask (generator.py:393,467), may be NULL
 *   out_bgr_u8 n*H*W*3 uint8 floor(clip(out,0,1)*255) BGR, may be NULL
 * Copies host->device, renders, copies device->host, returns when the outputs are valid. */
int rr_render_frames(rr_context *ctx, int n_frames, const uint8_t *bgr, const float *depth,
                     const rr_streak_rec *streaks, const int32_t *streak_offsets,
                     float *out_bgr, float *out_mask, uint8_t *out_bgr_u8);

/* Asynchronous form: rr_submit_frames enqueues the copies and kernels of one batch and returns; the
 * buffers of that batch (inputs and outputs, which must be page-locked for the copies to overlap)
 * belong to the library until the matching rr_wait_frames returns.  At most two batches are in flight
 * (a third submission first waits for the oldest); batches complete in submission order.  With two
 * sets of host buffers the host->device copy of batch k+1 and the device->host copy of batch k-1
 * overlap the kernels of batch k.  rr_render_frames == submit + wait. */
int rr_submit_frames(rr_context *ctx, int n_frames, const uint8_t *bgr, const float *depth,
                     const rr_streak_rec *streaks, const int32_t *streak_offsets,
                     float *out_bgr, float *out_mask, uint8_t *out_bgr_u8);
int rr_wait_frames(rr_context *ctx);

/* Frame formats at the boundary.  The float32 forms above are what the reference holds in memory; the compact forms
 * are what its files hold, and cut the bytes staged per frame from 14 to 9 per pixel:
 *   depth  RR_DEPTH_U16_256: the 16-bit samples of the depth PNG as cv2.imread(IMREAD_UNCHANGED) returns them; the
 *          library forms sample.astype(float32) / 256 (generator.py:360-365) on the device, exactly.
 *   mask   out_mask_idx8: what plt.imsave(path, rainy_mask) (generator.py:467) stores per pixel before the colour
 *          table: the float64 mask min/max-normalised, index = min(int(t * 256), 255) (matplotlib Normalize +
 *          Colormap.__call__; all zeros for a flat mask); out_mask_u16: int(t * 65535 + 0.5); out_mask_range:
 *          (min, max) of the float64 mask per frame, so that a caller can undo the normalisation.
 * Every output pointer may be NULL; outputs that are NULL are neither produced nor copied. */
enum { RR_DEPTH_F32_M = 0, RR_DEPTH_U16_256 = 1 };
typedef struct rr_frame_io {
    const uint8_t *bgr;              /* n*(rs*H)*(rs*W)*3 uint8                                   */
    const void *depth;               /* n*H*W float32 metres, or uint16 samples (depth_format)    */
    int32_t depth_format;            /* RR_DEPTH_F32_M / RR_DEPTH_U16_256                         */
    int32_t reserved;
    const rr_streak_rec *streaks;
    const int32_t *streak_offsets;   /* n+1, always a HOST array                                  */
    float *out_bgr;                  /* n*H*W*3 float32                                           */
    float *out_mask;                 /* n*H*W   float32                                           */
    uint8_t *out_bgr_u8;             /* n*H*W*3 uint8                                             */
    uint8_t *out_mask_idx8;          /* n*H*W   uint8                                             */
    uint16_t *out_mask_u16;          /* n*H*W   uint16                                            */
    double *out_mask_range;          /* n*2     float64 (min, max)                                */
    /* The saved files' image data finished on the device (host requests only; HOST buffers): complete zlib streams of the
     * Sub-filtered 8-bit RGBA scanlines -- what goes between "IDAT" and its CRC -- of the rainy image (from the uint8
     * output, alpha 255) and of the rain mask (its colormap index through matplotlib's viridis), as plt.imsave stores them
     * (generator.py:466-467).  Stream i starts at out_png_* + i * png_stride and is out_png_*_sizes[i] bytes long;
     * rr_host_png_write_streams frames and writes them.  png_stride >= rr_png_stream_bound(W, H). */
    uint8_t *out_png_image;
    uint8_t *out_png_mask;
    uint32_t *out_png_image_sizes;   /* n */
    uint32_t *out_png_mask_sizes;    /* n */
    size_t png_stride;
} rr_frame_io;
/* Upper bound of one GPU-made PNG stream of a W x H RGBA image (the device reserves the same). */
size_t rr_png_stream_bound(int W, int H);
int rr_render_frames_io(rr_context *ctx, int n_frames, const rr_frame_io *io);
int rr_submit_frames_io(rr_context *ctx, int n_frames, const rr_frame_io *io);     /* + rr_wait_frames */
/* DEVICE pointers in *io (streak_offsets stays on the host), no copies; asynchronous unless sync != 0. */
int rr_render_frames_device_io(rr_context *ctx, int n_frames, const rr_frame_io *io, int sync);

/* Same with DEVICE pointers and no copies (inputs already resident in HBM); asynchronous on the
 * context stream unless sync != 0. */
int rr_render_frames_device(rr_context *ctx, int n_frames, const uint8_t *d_bgr, const float *d_depth,
                            const rr_streak_rec *d_streaks, const int32_t *h_streak_offsets,
                            float *d_out_bgr, float *d_out_mask, uint8_t *d_out_bgr_u8, int sync);

/* Stage-level entry points for parity tests against oracle intermediates (host buffers). */
int rr_fog_only(rr_context *ctx, int n_frames, const uint8_t *bgr, const float *depth, double *out_planar_bgr_f64);
int rr_envmap_only(rr_context *ctx, int n_frames, const double *planar_bgr_f64, uint8_t *out_env_bgr_u8);
int rr_streak_photometry_only(rr_context *ctx, const uint8_t *env_bgr_u8, int n_streaks,
                              const rr_streak_rec *streaks, double *out_fovx_fovy_dropY /* n*3 */);

/* solid_angle.get_solid_angles (common/solid_angle.py:5-29) for an arbitrary (H_env, W_env) lat-long
 * grid, computed on the device; out is a HOST buffer of H_env*W_env float64. */
int rr_solid_angles(rr_context *ctx, int H_env, int W_env, double *out);

/* ---- on-the-fly particle simulation --------------------------------
 * B200-native stand-in for the closed AHLSimulation binary the reference drives through pexpect
 * (tools/simulation.py:259-469): Marshall-Palmer drop sizes, the binary's drag / gravity integrator
 * (semi-implicit Euler at sim_hz with the raindrop drag law recovered from its rodata), pinhole
 * shutter-open / shutter-close imaging and field-of-view culling.  Statistical parity only: the
 * binary cannot run here and its RNG is not reproducible.  One call produces the
 * imaged streaks of n_frames camera frames with the attributes of the simulator's XML <r> element
 * (common/bad_weather.py:200-211 reads exactly these). */
typedef struct rr_sim_params {
    int32_t W, H;            /* sensor size in px (full resolution, before render_scale)  */
    double focal_m;          /* cam_focal / 1000                                           */
    double pix_size_m;       /* cam_CCD_pixsize * 1e-6                                     */
    double exposure_ms;      /* cam_exposure                                               */
    double fallrate_mmh;     /* rain intensity                                             */
    double sim_hz;           /* integrator frequency, 2000 (common/db.py:64)               */
    double cam_speed_kmh;    /* forward camera motion (sim_steps["cam_motion"])            */
    double z_near, z_far;    /* visibility range in metres                                 */
    double d_min_mm, d_max_mm; /* drop size limits, 0.1 .. 10                              */
    double min_width_px;     /* drops narrower than this everywhere in range are not sampled */
    uint64_t seed;
} rr_sim_params;

typedef struct rr_sim_streak {  /* one <r .../> of the simulator XML, 120 bytes               */
    double wp1[3], wp2[3];   /* world start / end (camera relative, z negative forward)    */
    double wd1, wd2;         /* diameter in metres                                         */
    double ip1[2], ip2[2];   /* image start / end in px, y UP (the loader flips it)        */
    double iw1, iw2;         /* image width in px                                          */
    int64_t pid;
} rr_sim_streak;

/* out: capacity max_per_frame records per frame; counts[f] = records produced for frame f (the
 * expected count is returned through expected_per_frame when not NULL).  Frame f of the call is
 * simulated for absolute frame index first_frame + f: the result depends only on (seed, index). */
int rr_simulate_particles(rr_context *ctx, const rr_sim_params *p, int64_t first_frame, int n_frames,
                          int max_per_frame, rr_sim_streak *out, int32_t *counts, double *expected_per_frame);

/* The same simulation kept on the device, all the way to the renderer's input: one launch simulates n_frames frames, turns
 * the imaged drops into rr_streak_rec with the loader's arithmetic (common/bad_weather.py:200-238: positions / render_scale,
 * y flip with the render height, rounding, ratio, length, type), applies the in-frame filter (common/generator.py:413-420) and
 * makes each frame's NumPy RNG draws (np.random.seed(first_frame + f); generator.py:318,136; bad_weather.py:252-264) on the
 * device.  *d_records receives a DEVICE pointer to the concatenated records (library-owned, valid until the next call on
 * this context); only the n_frames + 1 offsets come back to the host (h_offsets) -- both go straight into
 * rr_render_frames_device_io.  Records are bit-identical to rr_simulate_particles + the host loader + rr_host_assemble_batch.
 * noise_std * noise_scale must be 0 (the wind write-back couples frames, generator.py:152-161). */
int rr_simulate_records_device(rr_context *ctx, const rr_sim_params *p, int64_t first_frame, int n_frames, int render_scale,
                               const double *db_ratios, int n_ratios, double noise_std, double noise_scale,
                               rr_streak_rec **d_records, int32_t *h_offsets, double *expected_per_frame);

int rr_debug_read(rr_context *ctx, int what, int frame, void *dst, size_t bytes);
/* Run-time options.  "serial" (0 / 1): 1 = the streak chain runs on the main stream, one kernel at a time (what a profiler sees;
 * rr_timings then gives every stage's own duration); 0 (default) = it runs beside the frame chain on a second stream. */
int rr_set_option(rr_context *ctx, const char *name, int value);
int rr_timings(rr_context *ctx, float *ms_per_stage /* RR_T_COUNT */);
int rr_kernel_launches(rr_context *ctx, long long *count);   /* kernels launched since rr_create */
int rr_stream(rr_context *ctx, void **cuda_stream);

/* Pinned host staging buffers for the callers of rr_render_frames (async copies need them). */
int rr_host_alloc(void **ptr, size_t bytes);
int rr_host_free(void *ptr);
/* The same with cudaHostAllocWriteCombined when write_combined != 0 (buffers the host only writes: the input sets). */
int rr_host_alloc_flags(void **ptr, size_t bytes, int write_combined);
/* What the host link of this process sustains right now: concurrent page-locked host->device (direction & 1) and
 * device->host (direction & 2) copies of `bytes` each on two streams for about `seconds`; GB/s per direction.
 * bench.py and tools/pcie_ceiling.py run it on every rank at once so that end-to-end numbers read as a fraction of
 * the box's measured ceiling. */
int rr_host_link_probe(int device_id, size_t bytes, double seconds, int write_combined, int direction,
                       double *h2d_gbs, double *d2h_gbs);

/* Host logic (no GPU): the per-frame NumPy legacy RNG draws of the reference, bit-exact --
 * np.random.seed(seed); per streak randint(10*bucket, 10*bucket+10) (bad_weather.py:252-264) and,
 * for non-Big streaks, normal(0, noise_std) * noise_scale (generator.py:136). */
int rr_host_draw_randoms(uint32_t seed, int n, const uint8_t *types, const int32_t *buckets, double noise_std,
                         double noise_scale, uint8_t *tex_idx, double *noise_deg);
/* Host logic (no GPU): the records of a whole batch of image frames from their simulator frames in one call -- the
 * in-frame filter (generator.py:413-420), the texture bucket (bad_weather.py:250-265) and the RNG draws above; frame f
 * is seeded with seeds[f] (generator.py:318).  Output records in frame order, offsets[n_frames + 1]; src_index (nullable)
 * = index of each output record in its simulator frame.  The wind rotation (generator.py:149-161) is left to the caller:
 * it is the identity when noise_std or noise_scale is 0.  RR_ERR_CAPACITY when out_cap records do not suffice. */
int rr_host_assemble_batch(int n_frames, const rr_streak_rec *const *sim, const int32_t *n_sim, const uint32_t *seeds,
                           int W, int H, const double *db_ratios, int n_ratios, double noise_std, double noise_scale,
                           rr_streak_rec *out, int64_t out_cap, int32_t *offsets, int32_t *src_index);
/* Host logic (no GPU): the field-of-view polygon of ONE streak in environment-map pixels -- FovComputation.
 * compute_fov_plane_points (common/bad_weather.py:596-704) for the camera at the origin and N = 20 cone rays: the 20 rays,
 * their lat-long image points and the wrap splice (20 or 24 vertices, x then y, into xy[48]).  It is the same header code
 * (csrc/rr_streak_geom.h) k_plan compiles for the device, evaluated on the host for callers that want the polygon itself;
 * the render path never calls it.  *n_vertices = 0 where the reference's try block would fail ("Drop skipped", :699-704). */
int rr_host_fov_polygon(const rr_streak_rec *rec, double radius, double fov_deg, int rows, int cols, double *xy,
                        int32_t *n_vertices);
/* Host logic (no GPU): native loader of the particle simulator's XML output -- replaces
 * DBManager.load_streaks_from_xml (common/bad_weather.py:148-248).  The root's children are camera
 * frames (<i id t d rs>), their children imaged streaks (<r pid wp1 wd1 wp2 wd2 ip1 iw1 ip2 iw2/>).
 * Every streak becomes an rr_streak_rec exactly as the reference builds its Streak object (:200-236:
 * /render_scale, y flip with the image height H, z negation, max_width, ratio, half-even rounding,
 * length, type) with tex_idx = 0 and noise_deg = 0 (drawn per frame later); kept iff max_width >= 1 and
 * length >= 1 (:238).  Streaks keep XML order; a repeated pid replaces the earlier entry in place and a
 * repeated frame id the earlier frame (dict.update, :238,241).  Malformed files return RR_ERR_ARG with
 * the reference's advice in rr_last_error() (:184-187). */
typedef struct rr_xml_frame {   /* one <i> element, 32 bytes                                   */
    int32_t id;              /* frame id (key of DBManager.streaks_simulator)                 */
    int32_t exposure_t;      /* "t"                                                           */
    int32_t start_d;         /* "d"                                                           */
    int32_t streaks_count;   /* "rs": the simulator's own count (before the loader's filter)  */
    int64_t first, count;    /* records [first, first + count) of the concatenated array      */
} rr_xml_frame;
typedef struct rr_xml_particles rr_xml_particles;
int rr_host_load_particles_xml(const char *path, int render_scale, int W, int H, rr_xml_particles **out);
int rr_host_particles_info(const rr_xml_particles *p, int32_t *n_frames, int64_t *n_records);
int rr_host_particles_copy(const rr_xml_particles *p, rr_xml_frame *frames, rr_streak_rec *records);
void rr_host_free_particles(rr_xml_particles *p);
/* np.linalg.norm of n 2-vectors the way NumPy's BLAS evaluates it, sqrt(fma(y, y, x * x)): the loader's
 * direction norm (bad_weather.py:229), shared with the Python-side record builder of the on-the-fly simulator. */
void rr_host_norm2(int n, const double *x, const double *y, double *out);
/* Host logic (no GPU): native PNG codec of the frame pipeline, on a pool of threads, straight from / into the batch
 * buffers of rr_submit_frames.  Decoding reproduces cv2.imread(path) (BGR uint8, common/generator.py:352) and
 * cv2.imread(path, IMREAD_UNCHANGED).astype(float32) / 256 (depth, generator.py:360-365) for non-interlaced 8/16-bit
 * gray / RGB(A) files; encoding writes what the drop-in Generator saves (generator.py:466-467): the uint8 image as RGB
 * and the rain mask min/max-normalised to 16-bit gray.  status[i] tells the caller which frames need its fallback
 * decoder.  level = zlib level 0..9 (0 stores, 1 is Huffman-only deflate). */
int rr_host_png_info(const char *path, int32_t *w, int32_t *h, int32_t *channels, int32_t *bit_depth);
int rr_host_png_read_batch(int n, const char *const *image_paths, const char *const *depth_paths, uint8_t *bgr, int Wi, int Hi,
                           float *depth, int Wd, int Hd, int n_threads, int32_t *status);
int rr_host_png_write_batch(int n, const char *const *image_paths, const uint8_t *bgr, const char *const *mask_paths,
                            const float *mask, int W, int H, int level, int n_threads);
/* The same decode with the depth delivered as the file's uint16 samples (rr_frame_io RR_DEPTH_U16_256). */
int rr_host_png_read_batch_u16(int n, const char *const *image_paths, const char *const *depth_paths, uint8_t *bgr, int Wi, int Hi,
                               uint16_t *depth, int Wd, int Hd, int n_threads, int32_t *status);
/* The reference's own file formats (plt.imsave, generator.py:466-467): 8-bit RGBA for both files, the mask coloured through
 * matplotlib's viridis table from its colormap index (rr_frame_io.out_mask_idx8).  level 1 = the library's own run +
 * Huffman deflate encoder (csrc/rr_host_deflate.h), 0 stores, 2..9 zlib. */
int rr_host_png_write_batch_rgba(int n, const char *const *image_paths, const uint8_t *bgr, const char *const *mask_paths,
                                 const uint8_t *mask_idx8, int W, int H, int level, int n_threads);
/* Compact files: 8-bit RGB image and 16-bit gray mask from rr_frame_io.out_mask_u16. */
int rr_host_png_write_batch_u16(int n, const char *const *image_paths, const uint8_t *bgr, const char *const *mask_paths,
                                const uint16_t *mask_u16, int W, int H, int level, int n_threads);
/* Frames n finished zlib streams (rr_frame_io.out_png_*) as 8-bit RGBA PNG files: signature, IHDR, one IDAT with its
 * CRC-32, IEND.  Returns the number of files that could not be written. */
int rr_host_png_write_streams(int n, const char *const *paths, const uint8_t *streams, size_t stride, const uint32_t *sizes,
                              int W, int H, int n_threads);
/* Test hook of the decoder the PNG reader uses (csrc/rr_host_inflate.h): zlib stream -> exactly out_len bytes. */
int rr_host_zlib_decompress_fast(const uint8_t *z, size_t zlen, uint8_t *out, size_t out_len);
/* Test hook of the deflate encoder: data -> zlib stream. */
int rr_host_zlib_compress_fast(const uint8_t *data, size_t n, uint8_t *out, size_t cap, size_t *out_len);
/* The simulator's force model evaluated on the host (CPU test-suite): terminal velocity solving
 * m g = F_drag(v), the drag at that speed and the drop mass. */
void rr_host_sim_physics(double D_m, double *v_terminal, double *drag_at_vt, double *mass);
/* The OpenCV Gaussian kernels baked into the library (for the CPU test-suite). */
void rr_host_tables(double *k64_25, float *k32_25, int *k15_fixed);
int rr_synchronize(rr_context *ctx);

#ifdef __cplusplus
}
#endif
#endif /* RAIN_B200_H */
