// C ABI of librain_b200.so (include/rain_b200.h): context, device memory, stage orchestration.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include "rr_kernels.cuh"
#include "rr_png_gpu.cuh"

static thread_local char g_err[1024] = "";
static void set_err(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
#define CK(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            set_err("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));            \
            return RR_ERR_CUDA;                                                                        \
        }                                                                                              \
    } while (0)

#define RR_MAX_SUB 8

struct rr_context {
    int device = 0, n_sm = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[RR_T_COUNT + 2];        // start of every timing slot
    cudaEvent_t ev_end[RR_T_COUNT + 2];    // end of the slots that run beside others (the streak chain on its own stream)
    bool slot_has_end[RR_T_COUNT + 2], slot_side[RR_T_COUNT + 2];
    float last_ms[RR_T_COUNT];
    // GPU-side PNG image data (rr_frame_io.out_png_*): scratch shared by the two passes, streams per async slot and kind
    bool png_ready = false;
    rr_png_bufs png;                       // scratch + geometry; .stream / .sizes are filled per use
    uint8_t *d_png_stream[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};      // [slot][0 image, 1 mask]
    unsigned *d_png_sizes[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    cudaStream_t s_png = nullptr;          // fetches the finished streams (exact sizes) after a batch completes
    rr_frame_io call_io[2];
    void *scratch[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};       // grow-only device buffers of the on-the-fly simulator
    size_t scratch_bytes[5] = {0, 0, 0, 0, 0};                // the requests in flight (host pointers of the PNG outputs)
    bool serial = false;                   // RR_SERIAL=1: the streak chain runs on the main stream (profiling: one kernel at a time)
    cudaEvent_t ev_plan = nullptr, ev_fext = nullptr;
    bool fext_side = false;                // RR_FEXT_SIDE=1: the extinction plane on the side stream beside the statistics (measured: 3.83 -> 3.88 ms, off)
    long long launches = 0;
    // streak DB
    uint8_t *d_db = nullptr;
    size_t db_bytes = 0;
    int n_tex = 0, db_width = 0;
    int32_t *d_tex_off = nullptr, *d_tex_h = nullptr, *d_tex_poff = nullptr;
    uint8_t *d_dbp = nullptr;            // zero-bordered texture copies, rebuilt from d_db before the first render after the DB changed
    bool padded_valid = false;
    int max_tex_h = 0;
    std::vector<int32_t> tex_heights;    // layout of the current DB (a DB of the same layout is written in place)
    // camera
    bool have_cam = false;
    rr_camera cam;
    rr_cam_dev camd;
    rr_fog_consts fogc;
    // k_fog's tile load: TMA boxes of the reflect-padded extinction planes (RR_FOG_TMA=0 selects the register-staged form)
    bool fog_tma = true;
    CUtensorMap fog_map, fog_map_roll;     // boxes of 88 x 56 (k_fog) and 88 x 32 (k_fog_roll) floats of the padded planes
    float *d_fext_lut = nullptr;
    int max_batch = 0, H_env = 0, W_env = 0, cyl_w = 0;
    int32_t *d_env_src = nullptr;
    uint8_t *d_env_written = nullptr, *d_env_tile_hole = nullptr;
    double *d_omega = nullptr, *d_omega_pref = nullptr, *d_omega_total = nullptr;
    // per batch
    uint8_t *d_bgr = nullptr;
    double *d_bgf = nullptr;         // reduced float64 image (render_scale == 2)
    float *d_depth = nullptr;        // float32 metres, or the uint16 PNG samples in its first half (rr_frame_io.depth_format)
    rr_streak_rec *d_streaks = nullptr;
    int32_t *d_offsets = nullptr;
    int streak_cap = 0;
    rr_frame_bufs fb;
    std::vector<void *> owned;       // everything cudaMalloc'ed for the camera (freed on re-set / destroy)
    int last_n_streaks = 0;
    // copy/compute overlap inside rr_render_frames (pinned host buffers): sub-batches on three streams
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    // k_plan (per-streak geometry, needs only the records) runs beside the frame stages of the same (sub-)batch
    cudaStream_t s_plan = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaEvent_t ev_in[RR_MAX_SUB], ev_done[RR_MAX_SUB], ev_d2h[RR_MAX_SUB], ev_out;
    int32_t *d_sub_offsets = nullptr;    // [RR_MAX_SUB][max_batch + 1]
    long long scan_base[RR_MAX_SUB];     // element index (in units of 6 long long) of each sub-batch's scan block
    int sub_n[RR_MAX_SUB];
    int n_sub_last = 1;
    // asynchronous submissions (rr_submit_frames / rr_wait_frames): at most two batches in flight
    cudaEvent_t ev_call[2];
    int inflight = 0, parity = 0;        // parity: slot the next submission uses; oldest in flight = parity ^ (inflight == 2 ? 0 : 1)
    int call_S[2] = {1, 1}, call_F[2] = {0, 0}, call_n[2] = {0, 0}, call_multi[2] = {0, 0};
    long long call_scan_base[2][RR_MAX_SUB];
    int call_sub_n[2][RR_MAX_SUB];
    int32_t *h_sub_off = nullptr;        // pinned staging [2][RR_MAX_SUB][max_batch + 1]
    int *d_err2 = nullptr;               // [2] overflow flags, one per slot
    int sub_cap = 0;                     // records per sub-batch slot of d_streaks
};

// allocates (releasing what *p pointed to before, if the context owned it)
template <class T>
static cudaError_t dev_alloc(rr_context *c, T **p, size_t count) {
    if (*p)
        for (size_t i = 0; i < c->owned.size(); i++)
            if (c->owned[i] == (void *)*p) { cudaFree(*p); c->owned.erase(c->owned.begin() + i); break; }
    *p = nullptr;
    cudaError_t e = cudaMalloc((void **)p, count * sizeof(T) + 256);
    if (e == cudaSuccess) c->owned.push_back((void *)*p);
    return e;
}

static void free_camera(rr_context *c) {
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->s_h2d); cudaStreamSynchronize(c->s_d2h); cudaStreamSynchronize(c->s_plan); }
    c->inflight = 0; c->parity = 0; c->sub_cap = 0;
    if (c->h_sub_off) { cudaFreeHost(c->h_sub_off); c->h_sub_off = nullptr; }
    for (void *p : c->owned) cudaFree(p);
    c->owned.clear();
    c->have_cam = false;
    c->streak_cap = 0;
    memset(&c->fb, 0, sizeof(c->fb));
    c->d_env_src = nullptr; c->d_env_written = nullptr; c->d_env_tile_hole = nullptr; c->d_omega = c->d_omega_pref = c->d_omega_total = nullptr;
    c->d_bgr = nullptr; c->d_bgf = nullptr; c->d_depth = nullptr; c->d_streaks = nullptr; c->d_offsets = nullptr;
    c->d_sub_offsets = nullptr; c->d_err2 = nullptr; c->d_fext_lut = nullptr;
    c->png_ready = false;
    memset(&c->png, 0, sizeof(c->png));
    memset(c->d_png_stream, 0, sizeof(c->d_png_stream)); memset(c->d_png_sizes, 0, sizeof(c->d_png_sizes));
}

extern "C" {

// cuTensorMapEncodeTiled through the runtime's driver entry point query: no link-time dependency on libcuda
typedef CUresult (*rr_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                       const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static rr_encode_tiled_fn encode_tiled_fn() {
    static rr_encode_tiled_fn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (rr_encode_tiled_fn)p;
        cudaGetLastError();
    }
    return fn;
}

// 3-D float32 tensor map (x fastest) of `planes` planes of rows x pitch elements; one box = box_w x box_h x 1
static int make_plane_map(CUtensorMap *map, void *base, int pitch, int rows, int planes, int box_w, int box_h) {
    rr_encode_tiled_fn enc = encode_tiled_fn();
    if (!enc) { set_err("cuTensorMapEncodeTiled is not available from this driver"); return RR_ERR_CUDA; }
    const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * sizeof(float), (cuuint64_t)pitch * rows * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1u}, estr[3] = {1u, 1u, 1u};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_err("cuTensorMapEncodeTiled failed (%d) for a %d x %d x %d float32 tensor", (int)r, pitch, rows, planes); return RR_ERR_CUDA; }
    return RR_OK;
}

size_t rr_png_stream_bound(int W, int H) {
    const size_t n = (size_t)H * ((size_t)4 * W + 1);
    return ((n + n / 4 + 4096) + 3) & ~(size_t)3;
}

int rr_version(void) { return 200; }
int rr_sim_device_of(rr_context *c) { return c ? c->device : 0; }
void *rr_ctx_stream(rr_context *c) { return c ? (void *)c->stream : nullptr; }
void rr_ctx_count_launches(rr_context *c, int n) { if (c) c->launches += n; }
void *rr_ctx_scratch(rr_context *c, int which, size_t bytes) {
    if (!c || which < 0 || which >= 5) return nullptr;
    if (bytes == 0) return c->scratch[which];                      // query: the buffer as it stands (NULL when never allocated)
    if (bytes > c->scratch_bytes[which]) {
        cudaStreamSynchronize(c->stream);
        if (c->scratch[which]) cudaFree(c->scratch[which]);
        c->scratch[which] = nullptr; c->scratch_bytes[which] = 0;
        const size_t want = bytes + bytes / 4 + 4096;
        if (cudaMalloc(&c->scratch[which], want) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        c->scratch_bytes[which] = want;
    }
    return c->scratch[which];
}
void rr_set_error(const char *msg) { set_err("%s", msg); }
const char *rr_last_error(void) { return g_err; }

int rr_create(int device_id, rr_context **out) {
    if (!out) { set_err("rr_create: out is NULL"); return RR_ERR_ARG; }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_err("rr_create: no usable CUDA device (%s); this library has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return RR_ERR_CUDA;
    }
    if (device_id < 0 || device_id >= n) { set_err("rr_create: device %d out of range (0..%d)", device_id, n - 1); return RR_ERR_ARG; }
    CK(cudaSetDevice(device_id));
    rr_context *c = new rr_context();
    c->device = device_id;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device_id));
    c->n_sm = prop.multiProcessorCount;
    // (a higher stream priority for the frame chain, the longer of the two chains of a step, was measured: 4.01 -> 4.13 ms --
    // the starved blur then delays the join; both chains run at the default priority)
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->s_plan, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->s_png, cudaStreamNonBlocking));
    memset(&c->png, 0, sizeof(c->png));
    memset(c->call_io, 0, sizeof(c->call_io));
    CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    for (int i = 0; i < RR_MAX_SUB; i++) {
        CK(cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c->ev_d2h[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 2; i++) CK(cudaEventCreateWithFlags(&c->ev_call[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_out, cudaEventDisableTiming));
    for (int i = 0; i < RR_T_COUNT + 2; i++) { CK(cudaEventCreate(&c->ev[i])); CK(cudaEventCreate(&c->ev_end[i])); c->slot_has_end[i] = c->slot_side[i] = false; }
    CK(cudaEventCreateWithFlags(&c->ev_plan, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_fext, cudaEventDisableTiming));
    { const char *env = getenv("RR_FEXT_SIDE"); c->fext_side = env && atoi(env) != 0; }
    { const char *env = getenv("RR_SERIAL"); c->serial = env && atoi(env) != 0; }
    memset(c->last_ms, 0, sizeof(c->last_ms));
    memset(&c->fb, 0, sizeof(c->fb));
    CK(rr_upload_constants());
    CK(rr_prepare_device());
    CK(rr_png_upload_constants());
    *out = c;
    return RR_OK;
}

int rr_destroy(rr_context *c) {
    if (!c) return RR_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    free_camera(c);
    for (int i = 0; i < 5; i++) if (c->scratch[i]) cudaFree(c->scratch[i]);
    if (c->d_db) cudaFree(c->d_db);
    if (c->d_tex_off) cudaFree(c->d_tex_off);
    if (c->d_tex_h) cudaFree(c->d_tex_h);
    if (c->d_tex_poff) cudaFree(c->d_tex_poff);
    if (c->d_dbp) cudaFree(c->d_dbp);
    for (int i = 0; i < RR_T_COUNT + 2; i++) { cudaEventDestroy(c->ev[i]); cudaEventDestroy(c->ev_end[i]); }
    cudaEventDestroy(c->ev_plan);
    cudaEventDestroy(c->ev_fext);
    cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->s_h2d);
    cudaStreamDestroy(c->s_d2h);
    cudaStreamSynchronize(c->s_plan);
    cudaStreamDestroy(c->s_plan);
    cudaStreamDestroy(c->s_png);
    cudaEventDestroy(c->ev_fork); cudaEventDestroy(c->ev_join);
    for (int i = 0; i < RR_MAX_SUB; i++) { cudaEventDestroy(c->ev_in[i]); cudaEventDestroy(c->ev_done[i]); cudaEventDestroy(c->ev_d2h[i]); }
    for (int i = 0; i < 2; i++) cudaEventDestroy(c->ev_call[i]);
    if (c->h_sub_off) cudaFreeHost(c->h_sub_off);
    cudaEventDestroy(c->ev_out);
    delete c;
    return RR_OK;
}

int rr_alloc_streak_db(rr_context *c, int n_tex, const int32_t *heights, int width) {
    if (!c || n_tex <= 0 || n_tex > 255 || !heights || width <= 0) { set_err("rr_alloc_streak_db: bad arguments"); return RR_ERR_ARG; }
    CK(cudaSetDevice(c->device));
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->s_plan); }
    if (c->d_db && n_tex == c->n_tex && width == c->db_width && (int)c->tex_heights.size() == n_tex &&
        memcmp(c->tex_heights.data(), heights, sizeof(int32_t) * n_tex) == 0) {
        c->padded_valid = false;         // same layout (a new Generator over the same streak DB): keep the buffers, new bytes follow
        return RR_OK;
    }
    if (c->d_db) { cudaFree(c->d_db); cudaFree(c->d_tex_off); cudaFree(c->d_tex_h); cudaFree(c->d_tex_poff); cudaFree(c->d_dbp); c->d_db = nullptr; }
    std::vector<int32_t> off(n_tex), poff(n_tex);
    size_t total = 0, ptotal = 0;
    c->max_tex_h = 0;
    for (int i = 0; i < n_tex; i++) {
        if (heights[i] <= 0) { set_err("rr_alloc_streak_db: texture %d has height %d", i, heights[i]); return RR_ERR_ARG; }
        off[i] = (int32_t)total;
        poff[i] = (int32_t)ptotal;
        total += (size_t)heights[i] * width;
        ptotal += (size_t)(heights[i] + 2) * (width + 2);
        if (heights[i] > c->max_tex_h) c->max_tex_h = heights[i];
    }
    if (ptotal > 0x7fffffff) { set_err("rr_alloc_streak_db: streak DB too large"); return RR_ERR_ARG; }
    CK(cudaMalloc((void **)&c->d_db, total + 256));
    CK(cudaMalloc((void **)&c->d_tex_off, sizeof(int32_t) * n_tex));
    CK(cudaMalloc((void **)&c->d_tex_h, sizeof(int32_t) * n_tex));
    CK(cudaMalloc((void **)&c->d_tex_poff, sizeof(int32_t) * n_tex));
    CK(cudaMalloc((void **)&c->d_dbp, ptotal + 256));
    CK(cudaMemcpy(c->d_tex_poff, poff.data(), sizeof(int32_t) * n_tex, cudaMemcpyHostToDevice));
    c->padded_valid = false;             // built from d_db right before the first render (the bytes may arrive by NCCL broadcast)
    CK(cudaMemcpy(c->d_tex_off, off.data(), sizeof(int32_t) * n_tex, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_tex_h, heights, sizeof(int32_t) * n_tex, cudaMemcpyHostToDevice));
    c->db_bytes = total; c->n_tex = n_tex; c->db_width = width;
    c->tex_heights.assign(heights, heights + n_tex);
    c->camd.db_width = width; c->camd.n_tex = n_tex;
    return RR_OK;
}

int rr_set_streak_db(rr_context *c, int n_tex, const int32_t *heights, int width, const uint8_t *gray_concat) {
    if (!gray_concat) { set_err("rr_set_streak_db: textures are NULL"); return RR_ERR_ARG; }
    int r = rr_alloc_streak_db(c, n_tex, heights, width);
    if (r != RR_OK) return r;
    CK(cudaMemcpy(c->d_db, gray_concat, c->db_bytes, cudaMemcpyHostToDevice));
    return RR_OK;
}

int rr_streak_db_device_ptr(rr_context *c, void **dev_ptr, size_t *bytes) {
    if (!c || !c->d_db) { set_err("rr_streak_db_device_ptr: no streak DB"); return RR_ERR_STATE; }
    if (dev_ptr) *dev_ptr = c->d_db;
    if (bytes) *bytes = c->db_bytes;
    c->padded_valid = false;             // the caller is about to write the textures (NCCL broadcast)
    return RR_OK;
}

static void drain(rr_context *c) {
    cudaStreamSynchronize(c->s_h2d); cudaStreamSynchronize(c->s_plan); cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->s_d2h);
}

// n: records of the whole batch (plans / sizes / boxes / scan); n_sub: records of the largest sub-batch slot
static int ensure_streak_cap(rr_context *c, int n, int n_sub = -1) {
    if (n_sub < 0) n_sub = n;
    if (n > c->streak_cap) {
        drain(c);
        int cap = n + n / 4 + 1024;
        CK(dev_alloc(c, &c->fb.plans, (size_t)cap));
        CK(cudaMemset(c->fb.plans, 0, sizeof(rr_plan) * (size_t)cap));     // the struct's padding is copied word by word (k_raster): keep initcheck quiet
        CK(dev_alloc(c, &c->fb.fcp, (size_t)cap));
        CK(dev_alloc(c, &c->fb.sizes, (size_t)cap));
        CK(dev_alloc(c, &c->fb.boxes, (size_t)cap));
        CK(dev_alloc(c, &c->fb.scan, (size_t)(cap + 1 + RR_MAX_SUB) * 6));
        c->streak_cap = cap;
    }
    if (n_sub > c->sub_cap) {
        drain(c);
        int cap = n_sub + n_sub / 4 + 1024;
        CK(dev_alloc(c, &c->d_streaks, (size_t)cap * RR_MAX_SUB));
        c->sub_cap = cap;
    }
    return RR_OK;
}

int rr_set_camera(rr_context *c, const rr_camera *cam, int max_batch) {
    if (!c || !cam || max_batch <= 0) { set_err("rr_set_camera: bad arguments"); return RR_ERR_ARG; }
    if (cam->render_scale != 0 && cam->render_scale != 1 && cam->render_scale != 2) { set_err("rr_set_camera: render_scale %d is not supported (1 or 2)", cam->render_scale); return RR_ERR_ARG; }
    if (cam->W < 32 || cam->H < 32 || cam->W > 8192 || cam->H > 8192) { set_err("rr_set_camera: unsupported size %dx%d", cam->W, cam->H); return RR_ERR_ARG; }
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    if (c->have_cam && max_batch == c->max_batch && memcmp(&c->cam, cam, sizeof(rr_camera)) == 0)
        return RR_OK;                    // same camera, same batch (a new Generator for the next sequence): tables and buffers stand
    free_camera(c);
    c->cam = *cam;
    c->max_batch = max_batch;
    const int W = cam->W, H = cam->H;
    // EnvironmentMapGenerator.__init__ / max_coord / min_coord  (bad_weather.py:708-740)
    int f = (int)(((cam->focal_m * 1000) / 12.7) * W);
    if (f <= 0) { set_err("rr_set_camera: focal length gives a %d px cylinder", f); return RR_ERR_ARG; }
    int cx = W / 2;
    int max_x = (int)rint(f * atan((double)cx / f) + cx);
    int min_x = (int)rint(f * atan((double)-cx / f) + cx);
    c->cyl_w = max_x - min_x + 1;
    c->H_env = H;
    c->W_env = c->cyl_w + 2 * (c->cyl_w / 2);
    const int He = c->H_env, We = c->W_env;
    if (We > 3072) { set_err("rr_set_camera: environment map width %d exceeds the supported 3072", We); return RR_ERR_ARG; }
    rr_cam_dev &d = c->camd;
    d.W = W; d.H = H; d.H_env = He; d.W_env = We;
    d.focal_m = cam->focal_m; d.f_number = cam->f_number; d.focus_plane = cam->focus_plane_m; d.pix_size = cam->pix_size_m;
    d.radius = cam->radius; d.fov_deg = cam->fov_deg; d.opacity_att = cam->opacity_att;
    d.exposure_blend = cam->exposure_ms / 1000.;                       // bad_weather.py:344
    d.db_width = c->db_width; d.n_tex = c->n_tex;
    // FogRain constants (add_attenuation.py:27-64)
    double beta_ext = 0.312 * pow(cam->fallrate_mmh, 0.67);
    c->fogc.neg_beta32 = (float)(-beta_ext);
    c->fogc.irr_scale_num = 4 * (cam->f_number * cam->f_number);
    c->fogc.irr_den = (cam->exposure_ms * 1e-3) * cam->gain * 3.141592653589793;
    {
        double g = 0.97;
        double cos_term = cos(90.0 * (3.141592653589793 / 180.0));   // math.cos(math.radians(90))
        c->fogc.beta_hg = (1 - (g * g)) / (4 * 3.141592653589793 * pow(1 + g * g - 2 * g * cos_term, 1.5));
    }
    // static tables
    CK(dev_alloc(c, &c->d_env_src, (size_t)He * We));
    CK(dev_alloc(c, &c->d_env_written, (size_t)He * We));
    CK(dev_alloc(c, &c->d_env_tile_hole, (size_t)((He + 15) / 16) * ((We + 63) / 64)));
    CK(dev_alloc(c, &c->d_omega, (size_t)He * We));
    CK(dev_alloc(c, &c->d_omega_pref, (size_t)He * (We + 1)));
    CK(dev_alloc(c, &c->d_omega_total, (size_t)1));
    {
        int32_t *scratch;
        CK(cudaMalloc((void **)&scratch, (size_t)H * c->cyl_w * 9 + 256));
        CK(rr_launch_env_tables(W, H, f, c->cyl_w, min_x, We, c->d_env_src, c->d_env_written, scratch, c->stream));
        CK(rr_launch_env_tile_flags(c->d_env_written, c->d_env_tile_hole, He, We, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaFree(scratch));
        c->launches += 5;
    }
    CK(rr_launch_omega(He, We, c->d_omega, c->d_omega_pref, c->d_omega_total, c->stream));
    c->launches += 3;
    // per-batch buffers
    const size_t F = (size_t)max_batch, np = (size_t)W * H;
    const int rs = cam->render_scale == 2 ? 2 : 1;
    rr_frame_bufs &b = c->fb;
    CK(dev_alloc(c, &c->d_bgr, F * np * 3 * rs * rs));
    CK(dev_alloc(c, &c->d_depth, F * np));
    CK(dev_alloc(c, &c->d_offsets, F + 1));
    CK(dev_alloc(c, &c->d_sub_offsets, (size_t)RR_MAX_SUB * (F + 1)));
    CK(cudaHostAlloc((void **)&c->h_sub_off, sizeof(int32_t) * 2 * RR_MAX_SUB * (F + 1), cudaHostAllocDefault));
    CK(dev_alloc(c, &c->d_err2, (size_t)2));
    CK(dev_alloc(c, &b.chan_sum, F * 4));
    CK(dev_alloc(c, &b.bg_sum, F * 4));
    CK(dev_alloc(c, &b.acs, F * 4));
    c->d_bgf = nullptr;
    if (rs == 2) CK(dev_alloc(c, &c->d_bgf, F * 3 * np));
    CK(dev_alloc(c, &b.rainy, F * 3 * np));
    CK(dev_alloc(c, &b.bg8, F * np * 4));
    {
        // extinction planes with the 12-pixel reflected halo materialised and a 16-byte pitch: every haloed fog tile is one TMA box
        const char *env = getenv("RR_FOG_TMA");
        c->fog_tma = !(env && atoi(env) == 0);
        const char *roll = getenv("RR_FOG_ROLL");
        b.fog_roll = c->fog_tma && !(roll && atoi(roll) == 0);
        b.fext_Wp = (W + 24 + 3) & ~3; b.fext_Hp = H + 24;
        CK(dev_alloc(c, &b.fext, F * (size_t)b.fext_Wp * b.fext_Hp));
        CK(dev_alloc(c, &c->d_fext_lut, (size_t)65536));
        CK(rr_launch_fext_lut(c->d_fext_lut, c->fogc.neg_beta32, c->stream));
        c->launches += 1;
        b.fext_lut = c->d_fext_lut;
        if (c->fog_tma) {
            int r = make_plane_map(&c->fog_map, b.fext, b.fext_Wp, b.fext_Hp, max_batch, 88, 56);
            if (r != RR_OK) return r;
            r = make_plane_map(&c->fog_map_roll, b.fext, b.fext_Wp, b.fext_Hp, max_batch, 88, 32);
            if (r != RR_OK) return r;
        }
    }
    CK(dev_alloc(c, &b.fblur, F * np));
    b.env_pitch = (We + 3) & ~3;
    {
        const char *env = getenv("RR_ENV_BULK");
        b.env_bulk = !(env && atoi(env) == 0);
    }
    CK(dev_alloc(c, &b.env8, F * (size_t)He * b.env_pitch * 4));
    CK(dev_alloc(c, &b.pref, F * RR_PREF_N * (size_t)He * (We + 1)));
    CK(dev_alloc(c, &b.rowtot, F * He));
    CK(dev_alloc(c, &b.ambient, F));
    b.err_flag = nullptr;
    CK(dev_alloc(c, &b.tile_sum, F * rr_n_partials(W, H)));             // also holds the 64x4 partial sums of k_downscale2
    CK(dev_alloc(c, &b.tile_min, F * rr_n_partials(W, H)));
    CK(dev_alloc(c, &b.tile_max, F * rr_n_partials(W, H)));
    CK(dev_alloc(c, &b.frame_mean, F));
    CK(dev_alloc(c, &b.maskd, F * np));
    CK(dev_alloc(c, &b.mask_range, F * 2));
    CK(dev_alloc(c, &b.out_bgr, F * np * 3));
    CK(dev_alloc(c, &b.out_mask, F * np));
    CK(dev_alloc(c, &b.out_u8, F * np * 3));
    CK(dev_alloc(c, &b.out_idx8, F * np));
    CK(dev_alloc(c, &b.out_u16, F * np));
    {
        double mult = 6.0;
        const char *env = getenv("RR_ARENA_MULT");
        if (env) mult = atof(env);
        b.arena_cap = (long long)(mult * (double)(F * np));
        CK(dev_alloc(c, &b.arena, (size_t)b.arena_cap));
    }
    int r = ensure_streak_cap(c, max_batch * 1024, max_batch * 1024);
    if (r != RR_OK) return r;
    c->have_cam = true;
    return RR_OK;
}

int rr_env_size(rr_context *c, int *H_env, int *W_env) {
    if (!c || !c->have_cam) { set_err("rr_env_size: no camera"); return RR_ERR_STATE; }
    if (H_env) *H_env = c->H_env;
    if (W_env) *W_env = c->W_env;
    return RR_OK;
}

// device buffers of the GPU PNG path, allocated on first use (max_batch streams of each kind per async slot)
static int ensure_png(rr_context *c) {
    if (c->png_ready) return RR_OK;
    drain(c);
    rr_png_bufs &p = c->png;
    const size_t F = (size_t)c->max_batch;
    p.W = c->cam.W; p.H = c->cam.H;
    p.n = (size_t)p.H * ((size_t)4 * p.W + 1);
    if (p.n * 15 + 4096 >= 0xffffffffull || (size_t)4 * p.W + 64 > 40000) { set_err("GPU PNG encoding: %d x %d is too large for this path", p.W, p.H); return RR_ERR_ARG; }
    p.n_pad = (p.n + 127) & ~(size_t)127;
    p.cap = rr_png_stream_bound(p.W, p.H);
    p.nchunks = (int)((p.n + 127) / 128);
    CK(dev_alloc(c, &p.filt, F * p.n_pad));
    CK(dev_alloc(c, &p.hist, F * 256));
    CK(dev_alloc(c, &p.adler, F * 2));
    CK(dev_alloc(c, &p.codes, F * 257));
    CK(dev_alloc(c, &p.chunk_bits, F * (size_t)p.nchunks));
    for (int s = 0; s < 2; s++)
        for (int k = 0; k < 2; k++) {
            CK(dev_alloc(c, &c->d_png_stream[s][k], F * p.cap));
            CK(dev_alloc(c, &c->d_png_sizes[s][k], F));
        }
    c->png_ready = true;
    return RR_OK;
}

// view of the PNG buffers for frames [f0, ...) of async slot `slot`, kind 0 image / 1 mask
static rr_png_bufs png_view(const rr_context *c, int slot, int kind, int f0) {
    rr_png_bufs v = c->png;
    v.filt += (size_t)f0 * v.n_pad; v.hist += (size_t)f0 * 256; v.adler += (size_t)f0 * 2; v.codes += (size_t)f0 * 257;
    v.chunk_bits += (size_t)f0 * v.nchunks;
    v.stream = c->d_png_stream[slot][kind] + (size_t)f0 * v.cap;
    v.sizes = c->d_png_sizes[slot][kind] + f0;
    return v;
}

static rr_static_tabs tabs_of(rr_context *c) {
    rr_static_tabs t;
    t.env_src = c->d_env_src; t.env_written = c->d_env_written; t.env_tile_hole = c->d_env_tile_hole; t.omega = c->d_omega; t.omega_pref = c->d_omega_pref;
    t.omega_total = c->d_omega_total; t.db = c->d_db; t.tex_off = c->d_tex_off; t.tex_h = c->d_tex_h;
    t.dbp = c->d_dbp; t.tex_poff = c->d_tex_poff;
    return t;
}

static int check_ready(rr_context *c, int n_frames, const char *who) {
    if (!c) { set_err("%s: context is NULL", who); return RR_ERR_ARG; }
    if (!c->have_cam) { set_err("%s: rr_set_camera has not been called", who); return RR_ERR_STATE; }
    if (n_frames <= 0 || n_frames > c->max_batch) { set_err("%s: n_frames %d not in 1..%d", who, n_frames, c->max_batch); return RR_ERR_ARG; }
    return RR_OK;
}

// The device pipeline for a batch whose inputs are already in fb.bgr / fb.depth / fb.streaks / fb.offsets.
// Two chains that meet at the compositor:
//   frame chain  (main stream):  stats -> extinction -> fog -> environment map -> prefix sums -> per-streak photometry
//   streak chain (side stream):  plan -> arena scan -> patch rasteriser -> defocus blur
// The rasteriser and the blur need only the geometry k_plan derives from the streak records, the photometry needs the
// frame's environment map; nothing in one chain reads what the other writes (k_setup fills the tint fields of the plans,
// the streak chain reads their geometry fields), so the integer / load-bound streak kernels run beside the float64-bound
// frame kernels on the same SMs.
static int run_pipeline(rr_context *c, int F, int n_streaks, bool timed) {
    rr_frame_bufs &b = c->fb;
    const rr_static_tabs t = tabs_of(c);
    cudaStream_t st = c->stream, ss = c->serial ? c->stream : c->s_plan;
    const int W = c->cam.W, H = c->cam.H;
    c->camd.db_width = c->db_width; c->camd.n_tex = c->n_tex;
    for (int i = 0; i < RR_T_COUNT; i++) c->slot_has_end[i] = c->slot_side[i] = false;
    if (!c->padded_valid) {              // first render with this streak DB: the rasteriser's zero-bordered texture copies
        CK(rr_launch_build_padded(c->d_db, c->d_tex_off, c->d_tex_h, c->d_tex_poff, c->n_tex, c->db_width, c->max_tex_h, c->d_dbp, st));
        c->padded_valid = true;
        c->launches += 1;
    }
    // fork: everything this (sub-)batch depends on has been enqueued on st (inputs landed, the previous user of the plan /
    // walker buffers and of the patch arena is done)
    if (!c->serial) { CK(cudaEventRecord(c->ev_fork, st)); CK(cudaStreamWaitEvent(ss, c->ev_fork, 0)); }
    if (timed) CK(cudaEventRecord(c->ev[RR_T_FOG], st));
    const int rs = c->cam.render_scale == 2 ? 2 : 1;
    if (c->serial) {
        CK(rr_launch_plan(b, t, c->camd, n_streaks, ss));
    }
    // the extinction plane depends on the depth only: on the side stream, beside the channel statistics
    const bool fext_side = !c->serial && c->fog_tma && c->fext_side;
    if (fext_side) { CK(rr_launch_fext_pad(b, c->fogc, F, W, H, ss)); CK(cudaEventRecord(c->ev_fext, ss)); }
    CK(rr_launch_stats(b, F, W, H, rs, (double *)b.bgf, st));
    if (fext_side) CK(cudaStreamWaitEvent(st, c->ev_fext, 0));
    CK(rr_launch_fog(b, c->fogc, F, W, H, c->fog_tma ? &c->fog_map : nullptr, &c->fog_map_roll, fext_side, st));
    if (timed) CK(cudaEventRecord(c->ev[RR_T_ENV], st));
    CK(rr_launch_env(b, t, F, W, H, c->W_env, st));
    if (timed) CK(cudaEventRecord(c->ev[RR_T_SETUP], st));
    // ---- streak chain ----
    if (!c->serial) {
        CK(rr_launch_plan(b, t, c->camd, n_streaks, ss));
        CK(cudaEventRecord(c->ev_plan, ss));
    }
    if (!c->serial) {
        CK(rr_launch_scan(b, n_streaks, ss));
        if (timed) CK(cudaEventRecord(c->ev[RR_T_RASTER], ss));
        CK(rr_launch_raster(b, t, c->camd, n_streaks, c->n_sm, ss));
        if (timed) { CK(cudaEventRecord(c->ev_end[RR_T_RASTER], ss)); CK(cudaEventRecord(c->ev[RR_T_BLUR], ss)); }
        CK(rr_launch_blur(b, n_streaks, c->n_sm, ss));
        if (timed) CK(cudaEventRecord(c->ev_end[RR_T_BLUR], ss));
        c->slot_has_end[RR_T_RASTER] = c->slot_has_end[RR_T_BLUR] = true;
        c->slot_side[RR_T_RASTER] = c->slot_side[RR_T_BLUR] = true;
        CK(cudaEventRecord(c->ev_join, ss));
        // ---- frame chain, continued: photometry needs the walkers k_plan prepared ----
        CK(cudaStreamWaitEvent(st, c->ev_plan, 0));
        CK(rr_launch_setup(b, t, c->camd, F, n_streaks, st));
        if (timed) CK(cudaEventRecord(c->ev_end[RR_T_SETUP], st));
        c->slot_has_end[RR_T_SETUP] = true;
        CK(cudaStreamWaitEvent(st, c->ev_join, 0));
    } else {
        CK(rr_launch_setup(b, t, c->camd, F, n_streaks, st));
        CK(rr_launch_scan(b, n_streaks, st));
        if (timed) CK(cudaEventRecord(c->ev[RR_T_RASTER], st));
        CK(rr_launch_raster(b, t, c->camd, n_streaks, c->n_sm, st));
        if (timed) CK(cudaEventRecord(c->ev[RR_T_BLUR], st));
        CK(rr_launch_blur(b, n_streaks, c->n_sm, st));
    }
    if (timed) CK(cudaEventRecord(c->ev[RR_T_COMPOSITE], st));
    // (compositor and epilogue in frame groups, the epilogue of a group beside the compositor of the next: measured 3.85 -> 3.87 /
    // 3.92 / 4.03 ms for 2 / 4 / 8 groups -- both are memory heavy and do not fill each other's gaps; one launch each)
    CK(rr_launch_composite(b, c->camd, F, st));
    if (timed) CK(cudaEventRecord(c->ev[RR_T_EPILOGUE], st));
    CK(rr_launch_epilogue(b, F, W, H, st));
    if (timed) CK(cudaEventRecord(c->ev[RR_T_D2H], st));
    // stats 2, fog constants + extinction + fog 3, env map + prefix + ambient 3, plan + set-up 2, scan 1, raster + blur 2,
    // composite + frame mean 2, epilogue 1
    c->launches += 2 + 3 + (b.fog_roll ? 1 : 0) + 3 + (n_streaks ? 2 : 0) + 1 + (n_streaks ? 2 : 0) + 2 + 1;
    c->last_n_streaks = n_streaks;
    return RR_OK;
}

// shifted view of the batch buffers for frames [f0, f0 + ...) whose streaks start at record s0
static rr_frame_bufs sub_view(const rr_context *c, const rr_frame_bufs &b, int f0, int s0, long long scan_base) {
    rr_frame_bufs v = b;
    const size_t np = (size_t)c->cam.W * c->cam.H;
    const size_t tiles = rr_n_partials(c->cam.W, c->cam.H);
    const int rs2 = c->cam.render_scale == 2 ? 4 : 1;
    v.bgr += (size_t)f0 * np * 3 * rs2; v.depth = (const char *)v.depth + (size_t)f0 * np * (v.depth_u16 ? 2 : 4); v.streaks += s0;
    if (v.bgf) v.bgf += (size_t)f0 * 3 * np;
    v.bg_sum += (size_t)f0 * 4; v.acs += (size_t)f0 * 4;
    v.chan_sum += (size_t)f0 * 4; v.rainy += (size_t)f0 * 3 * np; v.bg8 += (size_t)f0 * np * 4; v.fblur += (size_t)f0 * np; v.frame0 += f0;
    v.env8 += (size_t)f0 * c->H_env * b.env_pitch * 4;
    v.pref += (size_t)f0 * RR_PREF_N * c->H_env * (c->W_env + 1); v.rowtot += (size_t)f0 * c->H_env; v.ambient += f0;
    v.plans += s0; v.fcp += s0; v.sizes += s0; v.boxes += s0; v.scan += (size_t)scan_base * 6;
    v.tile_sum += (size_t)f0 * tiles; v.tile_min += (size_t)f0 * tiles; v.tile_max += (size_t)f0 * tiles; v.frame_mean += f0;
    v.maskd += (size_t)f0 * np; v.mask_range += (size_t)f0 * 2;
    if (v.out_bgr) v.out_bgr += (size_t)f0 * np * 3;
    if (v.out_mask) v.out_mask += (size_t)f0 * np;
    if (v.out_u8) v.out_u8 += (size_t)f0 * np * 3;
    if (v.out_idx8) v.out_idx8 += (size_t)f0 * np;
    if (v.out_u16) v.out_u16 += (size_t)f0 * np;
    return v;
}

static int finish_timings(rr_context *c) {
    // ev[RR_T_H2D] .. ev[RR_T_TOTAL] were recorded in order; slot i = time from ev[i] to ev[i+1]
    for (int i = RR_T_H2D; i < RR_T_TOTAL; i++) {
        float ms = 0;
        // a slot with its own end event (the side-stream stages; the set-up, after which the main stream waits for the side
        // stream) uses it; the others end where the next main-stream slot starts
        int nx = i + 1;
        while (nx < RR_T_TOTAL && c->slot_side[nx]) nx++;
        cudaError_t e = cudaEventElapsedTime(&ms, c->ev[i], c->slot_has_end[i] ? c->ev_end[i] : c->ev[nx]);
        c->last_ms[i] = e == cudaSuccess ? ms : -1.f;
    }
    float tot = 0;
    cudaError_t e = cudaEventElapsedTime(&tot, c->ev[RR_T_H2D], c->ev[RR_T_TOTAL]);
    c->last_ms[RR_T_TOTAL] = e == cudaSuccess ? tot : -1.f;
    return RR_OK;
}

// Overflow flag of one submission slot (the stream work of that slot must be complete).
// -> RR_OK, or RR_ERR_CAPACITY; with grow (nothing else in flight) the arena is enlarged to fit and the
// message says "grown" so that the caller re-runs the batch.
static int check_flag_slot(rr_context *c, const int *d_flag, int n_sub, const long long *scan_base, const int *sub_n, bool grow) {
    int flag = 0;
    CK(cudaMemcpy(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost));
    if (!flag) return RR_OK;
    long long need = 0;
    for (int k = 0; k < n_sub; k++) {
        long long v = 0;
        CK(cudaMemcpy(&v, c->fb.scan + (size_t)(scan_base[k] + sub_n[k]) * 6, sizeof(long long), cudaMemcpyDeviceToHost));
        if (v > need) need = v;
    }
    if (grow) {
        long long cap = need + need / 4 + 4096;
        double *na = nullptr;
        cudaError_t e = cudaMalloc((void **)&na, (size_t)cap * sizeof(double));
        if (e == cudaSuccess) {
            for (auto &p : c->owned) if (p == (void *)c->fb.arena) { cudaFree(p); p = (void *)na; }
            c->fb.arena = na; c->fb.arena_cap = cap;
            set_err("patch arena grown to %lld float64 elements", cap);
            return RR_ERR_CAPACITY;
        }
        cudaGetLastError();
    }
    set_err("patch arena overflow: need %lld float64 elements, have %lld (RR_ARENA_MULT raises the initial size; a synchronous "
            "rr_render_frames call grows it automatically)", need, c->fb.arena_cap);
    return RR_ERR_CAPACITY;
}

static int check_flag(rr_context *c, bool grow) {          // single-stream entry points (device / stage variants)
    CK(cudaStreamSynchronize(c->stream));
    return check_flag_slot(c, c->fb.err_flag, c->n_sub_last, c->scan_base, c->sub_n, grow);
}

// ---- asynchronous submission ---------------------------------------------------------------------
static int wait_oldest(rr_context *c, bool grow) {
    if (c->inflight == 0) return RR_OK;
    const int slot = c->inflight == 2 ? c->parity : (c->parity ^ 1);
    CK(cudaEventSynchronize(c->ev_call[slot]));
    c->inflight--;
    {   // the finished PNG streams: their sizes are on the host now, so exactly those bytes are fetched
        const rr_frame_io &io = c->call_io[slot];
        const int F = c->call_F[slot];
        for (int kind = 0; kind < 2; kind++) {
            uint8_t *host = kind ? io.out_png_mask : io.out_png_image;
            const uint32_t *sizes = kind ? io.out_png_mask_sizes : io.out_png_image_sizes;
            if (!host) continue;
            for (int i = 0; i < F; i++) {
                if (sizes[i] == 0 || sizes[i] > io.png_stride) { set_err("GPU PNG stream %d of the batch: %u bytes do not fit the %llu-byte stride", i, sizes[i], (unsigned long long)io.png_stride); return RR_ERR_CAPACITY; }
                CK(cudaMemcpyAsync(host + (size_t)i * io.png_stride, c->d_png_stream[slot][kind] + (size_t)i * c->png.cap, sizes[i], cudaMemcpyDeviceToHost, c->s_png));
            }
        }
        if (io.out_png_image || io.out_png_mask) CK(cudaStreamSynchronize(c->s_png));
    }
    return check_flag_slot(c, c->d_err2 + slot, c->call_S[slot], c->call_scan_base[slot], c->call_sub_n[slot], grow && c->inflight == 0);
}

// the device outputs a request does not ask for are not produced (NULL in the kernel's view)
static void select_outputs(rr_frame_bufs &v, const rr_frame_bufs &all, const rr_frame_io &io) {
    v.out_bgr = io.out_bgr ? all.out_bgr : nullptr;
    v.out_mask = io.out_mask ? all.out_mask : nullptr;
    v.out_u8 = (io.out_bgr_u8 || io.out_png_image) ? all.out_u8 : nullptr;
    v.out_idx8 = (io.out_mask_idx8 || io.out_png_mask) ? all.out_idx8 : nullptr;
    v.out_u16 = io.out_mask_u16 ? all.out_u16 : nullptr;
}

static int submit_frames(rr_context *c, int F, const rr_frame_io &io, bool timed, int default_sub) {
    const uint8_t *bgr = io.bgr;
    const rr_streak_rec *streaks = io.streaks;
    const int32_t *streak_offsets = io.streak_offsets;
    const int du16 = io.depth_format == RR_DEPTH_U16_256;
    const size_t dsz = du16 ? 2 : 4;
    const int n_streaks = streak_offsets[F];
    const size_t np = (size_t)c->cam.W * c->cam.H;
    const size_t rs2 = c->cam.render_scale == 2 ? 4 : 1;
    cudaStream_t st = c->stream;
    // With page-locked caller buffers the copies run on their own streams and the batch is cut into sub-batches
    // whose H2D / compute / D2H overlap.  Device inputs are single-buffered per sub-batch slot, so the copy-in of
    // the next batch's slot j starts as soon as this batch's slot j has been consumed: 4 slots for the
    // synchronous rr_render_frames (hides most of its own fill/drain), 2 for rr_submit_frames (larger kernels;
    // measured 7.3 k frames/s against 6.9 k with 4 and 5.1 k with 1 on the C2 benchmark).
    int S = 1;
    bool multi = false;
    {
        cudaPointerAttributes pa;
        multi = cudaPointerGetAttributes(&pa, bgr) == cudaSuccess && pa.type == cudaMemoryTypeHost;
        cudaGetLastError();
        const char *env = getenv(timed ? "RR_SUB_BATCHES" : "RR_SUB_BATCHES_ASYNC");
        int want = env ? atoi(env) : default_sub;
        if (multi && want > 1) S = want < RR_MAX_SUB ? want : RR_MAX_SUB;
        if (S > F) S = F;
    }
    int fstart[RR_MAX_SUB + 1], max_ns = 0;
    for (int k = 0; k <= S; k++) fstart[k] = (int)((long long)F * k / S);
    for (int k = 0; k < S; k++) { int ns = streak_offsets[fstart[k + 1]] - streak_offsets[fstart[k]]; if (ns > max_ns) max_ns = ns; }
    int r = ensure_streak_cap(c, n_streaks, max_ns);
    if (r != RR_OK) return r;
    if (io.out_png_image || io.out_png_mask) {
        if ((io.out_png_image && !io.out_png_image_sizes) || (io.out_png_mask && !io.out_png_mask_sizes) || io.png_stride < 64) {
            set_err("rr_frame_io: PNG outputs need their size arrays and a stride"); return RR_ERR_ARG;
        }
        r = ensure_png(c);
        if (r != RR_OK) return r;
    }
    if (c->inflight == 2) { r = wait_oldest(c, false); if (r != RR_OK) return r; }
    const int slot = c->parity, prev = slot ^ 1;
    if (c->inflight && (c->call_S[prev] != S || c->call_F[prev] != F || c->call_multi[prev] != (int)multi)) drain(c);    // slot regions would not line up
    rr_frame_bufs &b = c->fb;
    b.bgf = c->d_bgf;
    b.err_flag = c->d_err2 + slot;
    int32_t *sub_off = c->h_sub_off + (size_t)slot * RR_MAX_SUB * (c->max_batch + 1);
    if (timed) CK(cudaEventRecord(c->ev[RR_T_H2D], st));
    CK(cudaMemsetAsync(b.err_flag, 0, sizeof(int), st));
    long long sb = 0;
    for (int k = 0; k < S; k++) {
        const int f0 = fstart[k], nf = fstart[k + 1] - f0, s0 = streak_offsets[f0], ns = streak_offsets[f0 + nf] - s0;
        int32_t *so = sub_off + (size_t)k * (c->max_batch + 1);
        for (int i = 0; i <= nf; i++) so[i] = streak_offsets[f0 + i] - s0;
        cudaStream_t hs = multi ? c->s_h2d : st;
        rr_streak_rec *d_slot = c->d_streaks + (size_t)k * c->sub_cap;
        uint8_t *d_bgr = c->d_bgr;
        char *d_depth = (char *)c->d_depth;
        if (multi) CK(cudaStreamWaitEvent(hs, c->ev_done[k], 0));     // the previous user of this slot no longer reads its inputs
        CK(cudaMemcpyAsync(d_bgr + (size_t)f0 * np * 3 * rs2, bgr + (size_t)f0 * np * 3 * rs2, (size_t)nf * np * 3 * rs2, cudaMemcpyHostToDevice, hs));
        CK(cudaMemcpyAsync(d_depth + (size_t)f0 * np * dsz, (const char *)io.depth + (size_t)f0 * np * dsz, (size_t)nf * np * dsz, cudaMemcpyHostToDevice, hs));
        if (ns) CK(cudaMemcpyAsync(d_slot, streaks + s0, (size_t)ns * sizeof(rr_streak_rec), cudaMemcpyHostToDevice, hs));
        CK(cudaMemcpyAsync(c->d_sub_offsets + (size_t)k * (c->max_batch + 1), so, (nf + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, hs));
        if (multi) {
            CK(cudaEventRecord(c->ev_in[k], hs));
            CK(cudaStreamWaitEvent(st, c->ev_in[k], 0));
            CK(cudaStreamWaitEvent(st, c->ev_d2h[k], 0));            // the previous user's outputs of this slot have left the device
        }
        c->scan_base[k] = sb; c->sub_n[k] = ns;
        c->call_scan_base[slot][k] = sb; c->call_sub_n[slot][k] = ns;
        rr_frame_bufs saved = c->fb;
        saved.bgr = d_bgr; saved.depth = d_depth; saved.depth_u16 = du16; saved.streaks = d_slot - s0;   // sub_view adds s0 back
        saved.offsets = c->d_sub_offsets + (size_t)k * (c->max_batch + 1);
        rr_frame_bufs sel = saved;
        select_outputs(sel, saved, io);
        c->fb = sub_view(c, sel, f0, s0, sb);
        c->fb.offsets = saved.offsets;
        r = run_pipeline(c, nf, ns, timed && S == 1);
        if (r == RR_OK && (io.out_png_image || io.out_png_mask)) {
            // the saved files' image data, finished on the device (csrc/rr_png_gpu.cu)
            if (io.out_png_image) { cudaError_t e = rr_launch_png_encode(png_view(c, slot, 0, f0), c->fb.out_u8, false, nf, st); if (e != cudaSuccess) { set_err("GPU PNG encode: %s", cudaGetErrorString(e)); r = RR_ERR_CUDA; } }
            if (r == RR_OK && io.out_png_mask) { cudaError_t e = rr_launch_png_encode(png_view(c, slot, 1, f0), c->fb.out_idx8, true, nf, st); if (e != cudaSuccess) { set_err("GPU PNG encode: %s", cudaGetErrorString(e)); r = RR_ERR_CUDA; } }
            c->launches += 5 * ((io.out_png_image ? 1 : 0) + (io.out_png_mask ? 1 : 0));
        }
        saved.bgr = nullptr; saved.depth = nullptr; saved.streaks = nullptr;
        c->fb = saved;
        if (r != RR_OK) return r;
        sb += ns + 1;
        cudaStream_t ds = multi ? c->s_d2h : st;
        if (multi) { CK(cudaEventRecord(c->ev_done[k], st)); CK(cudaStreamWaitEvent(ds, c->ev_done[k], 0)); }
        const size_t fo = (size_t)f0;
        if (io.out_bgr) CK(cudaMemcpyAsync(io.out_bgr + (size_t)f0 * np * 3, b.out_bgr + fo * np * 3, (size_t)nf * np * 3 * sizeof(float), cudaMemcpyDeviceToHost, ds));
        if (io.out_mask) CK(cudaMemcpyAsync(io.out_mask + (size_t)f0 * np, b.out_mask + fo * np, (size_t)nf * np * sizeof(float), cudaMemcpyDeviceToHost, ds));
        if (io.out_bgr_u8) CK(cudaMemcpyAsync(io.out_bgr_u8 + (size_t)f0 * np * 3, b.out_u8 + fo * np * 3, (size_t)nf * np * 3, cudaMemcpyDeviceToHost, ds));
        if (io.out_mask_idx8) CK(cudaMemcpyAsync(io.out_mask_idx8 + (size_t)f0 * np, b.out_idx8 + fo * np, (size_t)nf * np, cudaMemcpyDeviceToHost, ds));
        if (io.out_mask_u16) CK(cudaMemcpyAsync(io.out_mask_u16 + (size_t)f0 * np, b.out_u16 + fo * np, (size_t)nf * np * 2, cudaMemcpyDeviceToHost, ds));
        if (io.out_mask_range) CK(cudaMemcpyAsync(io.out_mask_range + (size_t)f0 * 2, b.mask_range + fo * 2, (size_t)nf * 2 * sizeof(double), cudaMemcpyDeviceToHost, ds));
        if (io.out_png_image) CK(cudaMemcpyAsync(io.out_png_image_sizes + f0, c->d_png_sizes[slot][0] + f0, (size_t)nf * sizeof(uint32_t), cudaMemcpyDeviceToHost, ds));
        if (io.out_png_mask) CK(cudaMemcpyAsync(io.out_png_mask_sizes + f0, c->d_png_sizes[slot][1] + f0, (size_t)nf * sizeof(uint32_t), cudaMemcpyDeviceToHost, ds));
        if (multi) CK(cudaEventRecord(c->ev_d2h[k], ds));
    }
    c->n_sub_last = S;
    c->last_n_streaks = n_streaks;
    c->call_S[slot] = S; c->call_F[slot] = F; c->call_n[slot] = n_streaks; c->call_multi[slot] = (int)multi;
    c->call_io[slot] = io;
    if (timed) {
        if (multi) { CK(cudaEventRecord(c->ev_out, c->s_d2h)); CK(cudaStreamWaitEvent(st, c->ev_out, 0)); }
        CK(cudaEventRecord(c->ev[RR_T_TOTAL], st));
    }
    CK(cudaEventRecord(c->ev_call[slot], multi ? c->s_d2h : st));
    c->inflight++;
    c->parity ^= 1;
    return RR_OK;
}

static int validate_batch(rr_context *c, int n_frames, const rr_frame_io *io, const char *who) {
    int r = check_ready(c, n_frames, who);
    if (r != RR_OK) return r;
    if (!io || !io->bgr || !io->depth || !io->streak_offsets) { set_err("%s: NULL input", who); return RR_ERR_ARG; }
    if (io->depth_format != RR_DEPTH_F32_M && io->depth_format != RR_DEPTH_U16_256) { set_err("%s: unknown depth format %d", who, io->depth_format); return RR_ERR_ARG; }
    if (!c->d_db) { set_err("%s: no streak DB", who); return RR_ERR_STATE; }
    const int n_streaks = io->streak_offsets[n_frames];
    if (io->streak_offsets[0] != 0 || n_streaks < 0 || (n_streaks > 0 && !io->streaks)) { set_err("%s: bad streak offsets", who); return RR_ERR_ARG; }
    for (int f = 0; f < n_frames; f++)
        if (io->streak_offsets[f + 1] < io->streak_offsets[f]) { set_err("%s: streak offsets not monotone", who); return RR_ERR_ARG; }
    return RR_OK;
}

static rr_frame_io legacy_io(const uint8_t *bgr, const float *depth, const rr_streak_rec *streaks, const int32_t *streak_offsets,
                             float *out_bgr, float *out_mask, uint8_t *out_bgr_u8) {
    rr_frame_io io;
    memset(&io, 0, sizeof(io));
    io.bgr = bgr; io.depth = depth; io.depth_format = RR_DEPTH_F32_M; io.streaks = streaks; io.streak_offsets = streak_offsets;
    io.out_bgr = out_bgr; io.out_mask = out_mask; io.out_bgr_u8 = out_bgr_u8;
    return io;
}

int rr_submit_frames_io(rr_context *c, int n_frames, const rr_frame_io *io) {
    int r = validate_batch(c, n_frames, io, "rr_submit_frames_io");
    if (r != RR_OK) return r;
    CK(cudaSetDevice(c->device));
    return submit_frames(c, n_frames, *io, false, 2);
}

int rr_submit_frames(rr_context *c, int n_frames, const uint8_t *bgr, const float *depth, const rr_streak_rec *streaks,
                     const int32_t *streak_offsets, float *out_bgr, float *out_mask, uint8_t *out_bgr_u8) {
    const rr_frame_io io = legacy_io(bgr, depth, streaks, streak_offsets, out_bgr, out_mask, out_bgr_u8);
    return rr_submit_frames_io(c, n_frames, &io);
}

int rr_wait_frames(rr_context *c) {
    if (!c) { set_err("rr_wait_frames: context is NULL"); return RR_ERR_ARG; }
    CK(cudaSetDevice(c->device));
    return wait_oldest(c, true);
}

int rr_render_frames_io(rr_context *c, int n_frames, const rr_frame_io *io) {
    int r = validate_batch(c, n_frames, io, "rr_render_frames_io");
    if (r != RR_OK) return r;
    CK(cudaSetDevice(c->device));
    while (c->inflight) { r = wait_oldest(c, false); if (r != RR_OK) return r; }
    for (int attempt = 0;; attempt++) {
        r = submit_frames(c, n_frames, *io, true, 4);
        if (r != RR_OK) return r;
        const int S = c->n_sub_last;
        r = wait_oldest(c, attempt == 0);
        if (r == RR_ERR_CAPACITY && attempt == 0 && strstr(g_err, "grown")) continue;   // re-run once with the larger arena
        if (S == 1) finish_timings(c);
        else {
            memset(c->last_ms, 0, sizeof(c->last_ms));
            float tot = 0;
            if (cudaEventElapsedTime(&tot, c->ev[RR_T_H2D], c->ev[RR_T_TOTAL]) == cudaSuccess) c->last_ms[RR_T_TOTAL] = tot;
        }
        return r;
    }
}

int rr_render_frames(rr_context *c, int n_frames, const uint8_t *bgr, const float *depth, const rr_streak_rec *streaks,
                     const int32_t *streak_offsets, float *out_bgr, float *out_mask, uint8_t *out_bgr_u8) {
    const rr_frame_io io = legacy_io(bgr, depth, streaks, streak_offsets, out_bgr, out_mask, out_bgr_u8);
    return rr_render_frames_io(c, n_frames, &io);
}

// inputs and outputs of *io are DEVICE pointers (streak_offsets stays a host array); outputs may be NULL
static int render_device(rr_context *c, int F, const rr_frame_io &io, int sync) {
    int r;
    const int n_streaks = io.streak_offsets[F];
    while (c->inflight) { r = wait_oldest(c, false); if (r != RR_OK) return r; }
    r = ensure_streak_cap(c, n_streaks, 0);
    if (r != RR_OK) return r;
    c->fb.err_flag = c->d_err2;
    cudaStream_t st = c->stream;
    const rr_frame_bufs b_saved = c->fb;
    rr_frame_bufs &b = c->fb;
    for (int attempt = 0;; attempt++) {
        CK(cudaEventRecord(c->ev[RR_T_H2D], st));
        if (attempt == 0) CK(cudaMemcpyAsync(c->d_offsets, io.streak_offsets, (F + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        b.bgr = io.bgr; b.depth = io.depth; b.depth_u16 = io.depth_format == RR_DEPTH_U16_256; b.streaks = io.streaks; b.offsets = c->d_offsets;
        b.bgf = c->d_bgf;
        b.out_bgr = io.out_bgr; b.out_mask = io.out_mask; b.out_u8 = io.out_bgr_u8; b.out_idx8 = io.out_mask_idx8; b.out_u16 = io.out_mask_u16;
        c->n_sub_last = 1; c->scan_base[0] = 0; c->sub_n[0] = n_streaks;
        CK(cudaMemsetAsync(b.err_flag, 0, sizeof(int), st));
        r = run_pipeline(c, F, n_streaks, true);
        if (r == RR_OK && io.out_mask_range)
            CK(cudaMemcpyAsync(io.out_mask_range, b.mask_range, (size_t)F * 2 * sizeof(double), cudaMemcpyDeviceToDevice, st));
        b.out_bgr = b_saved.out_bgr; b.out_mask = b_saved.out_mask; b.out_u8 = b_saved.out_u8; b.out_idx8 = b_saved.out_idx8; b.out_u16 = b_saved.out_u16;
        b.bgr = nullptr; b.depth = nullptr; b.streaks = nullptr;
        if (r != RR_OK) return r;
        CK(cudaEventRecord(c->ev[RR_T_TOTAL], st));
        if (!sync) return RR_OK;
        r = check_flag(c, attempt == 0);
        if (r == RR_ERR_CAPACITY && attempt == 0 && strstr(g_err, "grown")) continue;      // re-run once with the larger arena
        finish_timings(c);
        return r;
    }
}

int rr_render_frames_device_io(rr_context *c, int n_frames, const rr_frame_io *io, int sync) {
    int r = check_ready(c, n_frames, "rr_render_frames_device_io");
    if (r != RR_OK) return r;
    if (!io || !io->bgr || !io->depth || !io->streak_offsets) { set_err("rr_render_frames_device_io: NULL input"); return RR_ERR_ARG; }
    if (io->depth_format != RR_DEPTH_F32_M && io->depth_format != RR_DEPTH_U16_256) { set_err("rr_render_frames_device_io: unknown depth format %d", io->depth_format); return RR_ERR_ARG; }
    if (!c->d_db) { set_err("rr_render_frames_device_io: no streak DB"); return RR_ERR_STATE; }
    CK(cudaSetDevice(c->device));
    return render_device(c, n_frames, *io, sync);
}

int rr_render_frames_device(rr_context *c, int n_frames, const uint8_t *d_bgr, const float *d_depth,
                            const rr_streak_rec *d_streaks, const int32_t *h_streak_offsets, float *d_out_bgr,
                            float *d_out_mask, uint8_t *d_out_bgr_u8, int sync) {
    const rr_frame_io io = legacy_io(d_bgr, d_depth, d_streaks, h_streak_offsets, d_out_bgr, d_out_mask, d_out_bgr_u8);
    return rr_render_frames_device_io(c, n_frames, &io, sync);
}

int rr_fog_only(rr_context *c, int n_frames, const uint8_t *bgr, const float *depth, double *out_planar) {
    int r = check_ready(c, n_frames, "rr_fog_only");
    if (r != RR_OK) return r;
    CK(cudaSetDevice(c->device));
    while (c->inflight) { r = wait_oldest(c, false); if (r != RR_OK) return r; }
    const size_t np = (size_t)c->cam.W * c->cam.H, F = n_frames;
    cudaStream_t st = c->stream;
    rr_frame_bufs &b = c->fb;
    const size_t rs2 = c->cam.render_scale == 2 ? 4 : 1;
    CK(cudaMemcpyAsync(c->d_bgr, bgr, F * np * 3 * rs2, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->d_depth, depth, F * np * sizeof(float), cudaMemcpyHostToDevice, st));
    b.bgr = c->d_bgr; b.depth = c->d_depth; b.depth_u16 = 0; b.bgf = c->d_bgf;
    CK(rr_launch_stats(b, n_frames, c->cam.W, c->cam.H, rs2 == 4 ? 2 : 1, c->d_bgf, st));
    b.frame0 = 0;
    CK(rr_launch_fog(b, c->fogc, n_frames, c->cam.W, c->cam.H, c->fog_tma ? &c->fog_map : nullptr, &c->fog_map_roll, false, st));
    c->launches += 5;
    CK(cudaMemcpyAsync(out_planar, b.rainy, F * 3 * np * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return RR_OK;
}

int rr_envmap_only(rr_context *c, int n_frames, const double *planar, uint8_t *out_env) {
    int r = check_ready(c, n_frames, "rr_envmap_only");
    if (r != RR_OK) return r;
    CK(cudaSetDevice(c->device));
    while (c->inflight) { r = wait_oldest(c, false); if (r != RR_OK) return r; }
    const size_t np = (size_t)c->cam.W * c->cam.H, F = n_frames, npe = (size_t)c->H_env * c->W_env;
    cudaStream_t st = c->stream;
    rr_frame_bufs &b = c->fb;
    CK(cudaMemcpyAsync(b.rainy, planar, F * 3 * np * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(rr_launch_planar_to_bg8(b.rainy, b.bg8, n_frames, c->cam.W, c->cam.H, st));
    CK(rr_launch_env(b, tabs_of(c), n_frames, c->cam.W, c->cam.H, c->W_env, st));
    c->launches += 4;
    std::vector<uint8_t> tmp(F * npe * 4);                     // the device keeps (B, G, R, 0) words in rows of env_pitch pixels
    CK(cudaMemcpy2DAsync(tmp.data(), (size_t)c->W_env * 4, b.env8, (size_t)b.env_pitch * 4, (size_t)c->W_env * 4, F * c->H_env, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (size_t i = 0; i < F * npe; i++) { out_env[3 * i] = tmp[4 * i]; out_env[3 * i + 1] = tmp[4 * i + 1]; out_env[3 * i + 2] = tmp[4 * i + 2]; }
    return RR_OK;
}

int rr_streak_photometry_only(rr_context *c, const uint8_t *env_bgr_u8, int n_streaks, const rr_streak_rec *streaks,
                              double *out3) {
    int r = check_ready(c, 1, "rr_streak_photometry_only");
    if (r != RR_OK) return r;
    if (!c->d_db) { set_err("rr_streak_photometry_only: no streak DB"); return RR_ERR_STATE; }
    if (n_streaks <= 0) return RR_OK;
    CK(cudaSetDevice(c->device));
    while (c->inflight) { r = wait_oldest(c, false); if (r != RR_OK) return r; }
    r = ensure_streak_cap(c, n_streaks, n_streaks);
    if (r != RR_OK) return r;
    const size_t npe = (size_t)c->H_env * c->W_env;
    cudaStream_t st = c->stream;
    rr_frame_bufs &b = c->fb;
    std::vector<uint8_t> env4(npe * 4, 0);                     // the device keeps (B, G, R, 0) words
    for (size_t i = 0; i < npe; i++) { env4[4 * i] = env_bgr_u8[3 * i]; env4[4 * i + 1] = env_bgr_u8[3 * i + 1]; env4[4 * i + 2] = env_bgr_u8[3 * i + 2]; }
    CK(cudaMemcpy2DAsync(b.env8, (size_t)b.env_pitch * 4, env4.data(), (size_t)c->W_env * 4, (size_t)c->W_env * 4, c->H_env, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->d_streaks, streaks, (size_t)n_streaks * sizeof(rr_streak_rec), cudaMemcpyHostToDevice, st));
    int32_t off[2] = {0, n_streaks};
    CK(cudaMemcpyAsync(c->d_offsets, off, sizeof(off), cudaMemcpyHostToDevice, st));
    b.streaks = c->d_streaks; b.offsets = c->d_offsets;
    rr_static_tabs t = tabs_of(c);
    // prefix sums of the given map, then the per-streak set-up
    CK(rr_launch_env_prefix_only(b, t, 1, c->H_env, c->W_env, st));
    c->camd.db_width = c->db_width; c->camd.n_tex = c->n_tex;
    CK(rr_launch_plan(b, t, c->camd, n_streaks, st));
    CK(rr_launch_setup(b, t, c->camd, 1, n_streaks, st));
    c->launches += 4;
    std::vector<rr_plan> plans(n_streaks);
    CK(cudaMemcpyAsync(plans.data(), b.plans, sizeof(rr_plan) * n_streaks, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int i = 0; i < n_streaks; i++) {
        out3[3 * i] = plans[i].valid ? plans[i].fov_x : NAN;
        out3[3 * i + 1] = plans[i].valid ? plans[i].fov_y : NAN;
        out3[3 * i + 2] = plans[i].valid ? plans[i].drop_Y : NAN;
    }
    return RR_OK;
}

int rr_solid_angles(rr_context *c, int H_env, int W_env, double *out) {
    if (!c || !out || H_env <= 0 || W_env <= 0) { set_err("rr_solid_angles: bad arguments"); return RR_ERR_ARG; }
    CK(cudaSetDevice(c->device));
    double *om = nullptr, *pref = nullptr, *tot = nullptr;
    size_t n = (size_t)H_env * W_env;
    CK(cudaMalloc((void **)&om, n * sizeof(double)));
    CK(cudaMalloc((void **)&pref, (size_t)H_env * (W_env + 1) * sizeof(double)));
    CK(cudaMalloc((void **)&tot, sizeof(double)));
    cudaError_t e = rr_launch_omega(H_env, W_env, om, pref, tot, c->stream);
    if (e == cudaSuccess) e = cudaMemcpy(out, om, n * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(om); cudaFree(pref); cudaFree(tot);
    c->launches += 3;
    CK(e);
    return RR_OK;
}

int rr_debug_read(rr_context *c, int what, int frame, void *dst, size_t bytes) {
    if (!c || !c->have_cam || !dst) { set_err("rr_debug_read: bad arguments"); return RR_ERR_ARG; }
    if (frame < 0 || frame >= c->max_batch) { set_err("rr_debug_read: frame out of range"); return RR_ERR_ARG; }
    CK(cudaSetDevice(c->device));
    while (c->inflight) { int q = wait_oldest(c, false); if (q != RR_OK) return q; }
    const size_t np = (size_t)c->cam.W * c->cam.H, npe = (size_t)c->H_env * c->W_env;
    const rr_frame_bufs &b = c->fb;
    const void *src = nullptr;
    size_t avail = 0;
    switch (what) {
        case RR_DBG_FOG_F64:
        case RR_DBG_RAINY_F64: src = b.rainy + (size_t)frame * 3 * np; avail = 3 * np * sizeof(double); break;
        case RR_DBG_ENV_U8: {                                  // device words (B, G, R, 0) -> packed BGR for the caller
            std::vector<uint8_t> tmp(npe * 4);
            CK(cudaStreamSynchronize(c->stream));
            CK(cudaMemcpy2D(tmp.data(), (size_t)c->W_env * 4, b.env8 + (size_t)frame * c->H_env * b.env_pitch * 4, (size_t)b.env_pitch * 4,
                            (size_t)c->W_env * 4, c->H_env, cudaMemcpyDeviceToHost));
            uint8_t *o = (uint8_t *)dst;
            for (size_t i = 0; i < npe && 3 * i + 2 < bytes; i++) { o[3 * i] = tmp[4 * i]; o[3 * i + 1] = tmp[4 * i + 1]; o[3 * i + 2] = tmp[4 * i + 2]; }
            return RR_OK;
        }
        case RR_DBG_OMEGA: src = c->d_omega; avail = npe * sizeof(double); break;
        case RR_DBG_PLANS: src = b.plans; avail = sizeof(rr_plan) * (size_t)c->last_n_streaks; break;
        case RR_DBG_ENV_SRC: src = c->d_env_src; avail = npe * sizeof(int32_t); break;
        case RR_DBG_ARENA: src = b.arena; avail = (size_t)b.arena_cap * sizeof(double); break;
        case RR_DBG_FEXT: src = b.fblur + (size_t)frame * np; avail = np * sizeof(float); break;
        default: set_err("rr_debug_read: unknown selector %d", what); return RR_ERR_ARG;
    }
    if (bytes > avail) bytes = avail;
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return RR_OK;
}

int rr_set_option(rr_context *c, const char *name, int value) {
    if (!c || !name) { set_err("rr_set_option: bad arguments"); return RR_ERR_ARG; }
    CK(cudaSetDevice(c->device));
    while (c->inflight) { int q = wait_oldest(c, false); if (q != RR_OK) return q; }
    drain(c);
    if (!strcmp(name, "serial")) { c->serial = value != 0; return RR_OK; }
    set_err("rr_set_option: unknown option '%s'", name);
    return RR_ERR_ARG;
}

int rr_timings(rr_context *c, float *ms) {
    if (!c || !ms) { set_err("rr_timings: bad arguments"); return RR_ERR_ARG; }
    memcpy(ms, c->last_ms, sizeof(c->last_ms));
    return RR_OK;
}

int rr_kernel_launches(rr_context *c, long long *count) {
    if (!c || !count) { set_err("rr_kernel_launches: bad arguments"); return RR_ERR_ARG; }
    *count = c->launches;
    return RR_OK;
}

int rr_stream(rr_context *c, void **s) {
    if (!c || !s) { set_err("rr_stream: bad arguments"); return RR_ERR_ARG; }
    *s = (void *)c->stream;
    return RR_OK;
}

int rr_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) { set_err("rr_host_alloc: ptr is NULL"); return RR_ERR_ARG; }
    CK(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return RR_OK;
}

int rr_host_alloc_flags(void **ptr, size_t bytes, int write_combined) {
    if (!ptr) { set_err("rr_host_alloc_flags: ptr is NULL"); return RR_ERR_ARG; }
    CK(cudaHostAlloc(ptr, bytes ? bytes : 1, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
    return RR_OK;
}

// Concurrent pinned host->device and device->host copies on two streams for `seconds`: what the host link of this
// process (PCIe + host memory, under whatever the other ranks of the box are doing at the same time) sustains.
int rr_host_link_probe(int device_id, size_t bytes, double seconds, int write_combined, int direction, double *h2d_gbs,
                       double *d2h_gbs) {
    if (bytes < (1u << 20) || seconds <= 0 || !(direction & 3)) { set_err("rr_host_link_probe: bad arguments"); return RR_ERR_ARG; }
    CK(cudaSetDevice(device_id));
    void *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
    cudaStream_t s[2] = {nullptr, nullptr};
    cudaEvent_t e0[2] = {nullptr, nullptr}, e1[2] = {nullptr, nullptr};
    int rc = RR_OK;
    double gbs[2] = {0, 0};
    do {
        cudaError_t e;
#define PB(call) if ((e = (call)) != cudaSuccess) { set_err("rr_host_link_probe: %s -> %s", #call, cudaGetErrorString(e)); rc = RR_ERR_CUDA; break; }
        PB(cudaHostAlloc(&h_in, bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault));
        PB(cudaHostAlloc(&h_out, bytes, cudaHostAllocDefault));
        memset(h_in, 1, bytes); memset(h_out, 0, bytes);            // first touch on the calling thread's NUMA node
        PB(cudaMalloc(&d_in, bytes)); PB(cudaMalloc(&d_out, bytes));
        PB(cudaMemset(d_out, 2, bytes));
        for (int i = 0; i < 2; i++) { PB(cudaStreamCreateWithFlags(&s[i], cudaStreamNonBlocking)); PB(cudaEventCreate(&e0[i])); PB(cudaEventCreate(&e1[i])); }
        if (rc != RR_OK) break;
        // warm-up
        if (direction & 1) PB(cudaMemcpyAsync(d_in, h_in, bytes, cudaMemcpyHostToDevice, s[0]));
        if (direction & 2) PB(cudaMemcpyAsync(h_out, d_out, bytes, cudaMemcpyDeviceToHost, s[1]));
        PB(cudaStreamSynchronize(s[0])); PB(cudaStreamSynchronize(s[1]));
        // size the run from one timed round
        long long reps = 1, done = 0;
        double elapsed = 0;
        for (int i = 0; i < 2; i++) PB(cudaEventRecord(e0[i], s[i]));
        while (elapsed < seconds && rc == RR_OK) {
            for (long long k = 0; k < reps; k++) {
                if (direction & 1) PB(cudaMemcpyAsync(d_in, h_in, bytes, cudaMemcpyHostToDevice, s[0]));
                if (direction & 2) PB(cudaMemcpyAsync(h_out, d_out, bytes, cudaMemcpyDeviceToHost, s[1]));
            }
            for (int i = 0; i < 2; i++) PB(cudaEventRecord(e1[i], s[i]));
            PB(cudaStreamSynchronize(s[0])); PB(cudaStreamSynchronize(s[1]));
            done += reps;
            float ms0 = 0, ms1 = 0;
            PB(cudaEventElapsedTime(&ms0, e0[0], e1[0])); PB(cudaEventElapsedTime(&ms1, e0[1], e1[1]));
            elapsed = (ms0 > ms1 ? ms0 : ms1) / 1000.0;
            if (direction & 1) gbs[0] = (double)done * bytes / (ms0 / 1000.0) / 1e9;
            if (direction & 2) gbs[1] = (double)done * bytes / (ms1 / 1000.0) / 1e9;
        }
#undef PB
    } while (0);
    for (int i = 0; i < 2; i++) { if (e0[i]) cudaEventDestroy(e0[i]); if (e1[i]) cudaEventDestroy(e1[i]); if (s[i]) cudaStreamDestroy(s[i]); }
    if (d_in) cudaFree(d_in); if (d_out) cudaFree(d_out);
    if (h_in) cudaFreeHost(h_in); if (h_out) cudaFreeHost(h_out);
    if (h2d_gbs) *h2d_gbs = gbs[0];
    if (d2h_gbs) *d2h_gbs = gbs[1];
    return rc;
}

int rr_host_free(void *ptr) {
    if (ptr) CK(cudaFreeHost(ptr));
    return RR_OK;
}

int rr_synchronize(rr_context *c) {
    if (!c) { set_err("rr_synchronize: context is NULL"); return RR_ERR_ARG; }
    CK(cudaSetDevice(c->device));
    int r = RR_OK;
    while (c->inflight) { int q = wait_oldest(c, false); if (q != RR_OK) r = q; }
    CK(cudaStreamSynchronize(c->stream));
    if (c->have_cam && c->fb.err_flag) { int q = check_flag(c, false); if (q != RR_OK) r = q; }
    if (c->have_cam) finish_timings(c);
    return r;
}

}  // extern "C"
