// CUDA kernels of the rain-rendering hot path (sm_100a).  Compiled with -fmad=false: every
// float64 expression keeps the reference's operation order without FMA contraction.
// Reference citations are relative to astra-vision/rain-rendering.
#include "rr_kernels.cuh"
#include <stdio.h>

// ------------------------------------------------------------------------------------------
// TMA (cp.async.bulk[.tensor]) + mbarrier, inline PTX for sm_100a
// ------------------------------------------------------------------------------------------
// streaming (evict-first) store for data written once and read much later or never by this GPU: keeps it from pushing
// reusable lines out of L2.  RR_NO_STREAM_STORES=1 at build time turns them into plain stores (A/B measurements).
#ifdef RR_NO_STREAM_STORES
#define RR_STREAM_STORE(ptr, val) (*(ptr) = (val))
#else
#define RR_STREAM_STORE(ptr, val) __stcs((ptr), (val))
#endif

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");       // visible to the async proxy before a copy signals it
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "RR_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra RR_MBAR_DONE;\n"
        "bra RR_MBAR_WAIT;\n"
        "RR_MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one box of a 3-D tensor map -> shared memory; completion (box bytes) is signalled on the mbarrier
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int x, int y, int z, unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
                 : "memory");
}
// bytes (a multiple of 16, both addresses 16-byte aligned) global -> shared memory
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------------------------------
// constants (uploaded once by rr_upload_constants)
// ------------------------------------------------------------------------------------------
__constant__ double c_k64[25];      // cv2.getGaussianKernel(25, 25, CV_64F)   (add_attenuation.py:79-80)
__constant__ double c_k32d[25];     // cv2.getGaussianKernel(25, 25, CV_32F), the float32 values widened to float64
__constant__ int c_k15[15];         // OpenCV fixed-point (8 fractional bits) kernel of GaussianBlur((15,15), 0) on uint8
__device__ float d_cubic[RR_INTER_TAB * 4];   // divergent per-thread indexing: global/L1, not the constant bank

// half of the symmetric 25-tap kernel as OpenCV 4.13 computes it (bit-exact softfloat path);
// tests/test_host_logic.py checks these against cv2.getGaussianKernel.
static const double h_k64_half[13] = {
    0x1.30388bb7cc924p-5, 0x1.35ded00af4b8ep-5, 0x1.3b1ec2bb8377ap-5, 0x1.3ff2529db4fc4p-5, 0x1.4453db9cbbcb7p-5,
    0x1.483e31b371756p-5, 0x1.4bacab167051ep-5, 0x1.4e9b2973a742ep-5, 0x1.5106222dbfd23p-5, 0x1.52eaa57c51c85p-5,
    0x1.5446645cd4cffp-5, 0x1.5517b5437ec3fp-5, 0x1.555d977eb8796p-5};
static const int h_k15[15] = {1, 3, 6, 12, 20, 30, 36, 40, 36, 30, 20, 12, 6, 3, 1};

extern "C" void rr_host_tables(double *k64, float *k32, int *k15) {
    for (int i = 0; i < 25; i++) {
        double v = h_k64_half[i <= 12 ? i : 24 - i];
        k64[i] = v;
        k32[i] = (float)v;
    }
    for (int i = 0; i < 15; i++) k15[i] = h_k15[i];
}

cudaError_t rr_upload_constants() {
    double k64[25]; float k32[25]; int k15[15];
    rr_host_tables(k64, k32, k15);
    float cub[RR_INTER_TAB * 4];
    rr_build_cubic_tab(cub);
    cudaError_t e;
    if ((e = cudaMemcpyToSymbol(c_k64, k64, sizeof(k64))) != cudaSuccess) return e;
    double k32d[25];
    for (int i = 0; i < 25; i++) k32d[i] = (double)k32[i];
    if ((e = cudaMemcpyToSymbol(c_k32d, k32d, sizeof(k32d))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(c_k15, k15, sizeof(k15))) != cudaSuccess) return e;
    if ((e = cudaMemcpyToSymbol(d_cubic, cub, sizeof(cub))) != cudaSuccess) return e;
    return cudaSuccess;
}

__device__ __forceinline__ int r101(int i, int n) {   // BORDER_REFLECT_101
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------
// init-time: cylindrical environment-map tables  (bad_weather.py:742-812, geometry only)
// ------------------------------------------------------------------------------------------
__global__ void k_cyl_scatter(int W, int H, int f, int min_x, int cyl_w, int32_t *first) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= W * H) return;
    int yy = idx / W, xx = idx - yy * W;
    int cx = W / 2, cy = H / 2;
    double hh = (double)xx - cx, vv = (double)yy - cy;
    double fd = (double)f;
    double rowp = rint((fd * (vv / sqrt(hh * hh + (double)(f * f)))) + cy);      // :724-725,760
    double colp = rint((fd * atan(hh / fd)) + cx) - min_x;                       // :726,760-761
    int r = (int)rowp, c = (int)colp;
    if (r < 0 || r >= H || c < 0 || c >= cyl_w) return;
    atomicMin(&first[r * cyl_w + c], idx);    // np.unique(..., return_index=True): lowest flat index wins (:762-767)
}

__global__ void k_cyl_fill(int H, int cyl_w, const int32_t *first, int32_t *filled, uint8_t *written) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cyl_w) return;
    const int NONE = 0x7fffffff;
    int half = H / 2;
    int y_up = 0;
    for (int y = 0; y < half; y++) if (first[y * cyl_w + c] != NONE) { y_up = y; break; }     // np.argmax(mask>0) (:827)
    int r_dn = 0;
    for (int r = 0; r < H - half; r++) if (first[(H - 1 - r) * cyl_w + c] != NONE) { r_dn = r; break; }   // :839-841
    int src_up = first[y_up * cyl_w + c], src_dn = first[(H - 1 - r_dn) * cyl_w + c];
    for (int y = 0; y < H; y++) {
        int v = first[y * cyl_w + c];
        bool w = v != NONE;
        if (!w) {
            if (y < half) v = src_up;                  // :785-789
            else if (y >= H - half) v = src_dn;        // :776-781
        }
        filled[y * cyl_w + c] = (v == NONE) ? -1 : v;
        written[y * cyl_w + c] = w ? 1 : 0;
    }
}

__global__ void k_cyl_expand(int H, int cyl_w, int W_env, const int32_t *filled, const uint8_t *written,
                             int32_t *env_src, uint8_t *env_written) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * W_env) return;
    int y = i / W_env, j = i - y * W_env;
    int pad = cyl_w / 2;
    int len = cyl_w - cyl_w / 2, start = W_env - len;
    int c;
    if (j >= start) c = cyl_w - 1 - (j - start);       // :806-812
    else if (j < pad) c = pad - 1 - j;                 // :797-803
    else c = j - pad;                                  // :791
    env_src[i] = filled[y * cyl_w + c];
    env_written[i] = written[y * cyl_w + c];
}

__global__ void k_fill_i32(int32_t *p, int32_t v, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

cudaError_t rr_launch_env_tables(int W, int H, int focal_px, int cyl_w, int min_x, int W_env, int32_t *env_src,
                                 uint8_t *env_written, int32_t *scratch, cudaStream_t st) {
    // scratch: [H*cyl_w] first, [H*cyl_w] filled, then H*cyl_w bytes written
    int32_t *first = scratch, *filled = scratch + (size_t)H * cyl_w;
    uint8_t *written = (uint8_t *)(filled + (size_t)H * cyl_w);
    k_fill_i32<<<(unsigned)(((size_t)H * cyl_w + 255) / 256), 256, 0, st>>>(first, 0x7fffffff, (size_t)H * cyl_w);
    k_cyl_scatter<<<(W * H + 255) / 256, 256, 0, st>>>(W, H, focal_px, min_x, cyl_w, first);
    k_cyl_fill<<<(cyl_w + 127) / 128, 128, 0, st>>>(H, cyl_w, first, filled, written);
    k_cyl_expand<<<(H * W_env + 255) / 256, 256, 0, st>>>(H, cyl_w, W_env, filled, written, env_src, env_written);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// init-time: per-pixel solid angles of the lat-long map  (solid_angle.py:5-102)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void sa_corner(int ci, int ri, int W_env, int H_env, double v[3]) {
    // np.linspace(0, 1, n + 1): i * (1/n), last element exactly 1
    double u = (ci == W_env) ? 1.0 : ci * (1.0 / W_env);
    double w = (ri == H_env) ? 1.0 : ri * (1.0 / H_env);
    u = u * 2;
    double theta = RR_PI * (u - 1);
    double phi = RR_PI * w;
    double sp = sin(phi);
    v[0] = sp * sin(theta);
    v[1] = cos(phi);
    v[2] = -sp * cos(theta);
}

__device__ __forceinline__ double sa_tetra(const double a[3], const double b[3], const double c[3]) {
    double ta = acos((b[0] * c[0] + b[1] * c[1]) + b[2] * c[2]);
    double tb = acos((a[0] * c[0] + a[1] * c[1]) + a[2] * c[2]);
    double tc = acos((a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]);
    double ts = (ta + tb + tc) / 2;
    double product = tan(ts / 2) * tan((ts - ta) / 2) * tan((ts - tb) / 2) * tan((ts - tc) / 2);
    if (product < 0) product = 0;
    return 4 * atan(sqrt(product));
}

__global__ void k_omega(int H_env, int W_env, double *omega) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H_env * W_env) return;
    int r = i / W_env, c = i - r * W_env;
    double a[3], b[3], cc[3], d[3];
    sa_corner(c, r, W_env, H_env, a);
    sa_corner(c + 1, r, W_env, H_env, b);
    sa_corner(c, r + 1, W_env, H_env, cc);
    sa_corner(c + 1, r + 1, W_env, H_env, d);
    double o = sa_tetra(a, b, cc);
    o += sa_tetra(b, cc, d);
    omega[i] = o;
}

// one block per row: exclusive prefix (W_env + 1 entries) and the row total
__global__ void k_row_prefix1(int W_env, const double *vals, double *pref, double *rowtot) {
    int r = blockIdx.x;
    if (threadIdx.x == 0) {
        const double *v = vals + (size_t)r * W_env;
        double *p = pref + (size_t)r * (W_env + 1);
        double s = 0;
        for (int c = 0; c < W_env; c++) { p[c] = s; s += v[c]; }
        p[W_env] = s;
        rowtot[r] = s;
    }
}

__global__ void k_sum_serial(int n, const double *v, double *out) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double s = 0;
        for (int i = 0; i < n; i++) s += v[i];
        *out = s;
    }
}

cudaError_t rr_launch_omega(int H_env, int W_env, double *omega, double *omega_pref, double *omega_total, cudaStream_t st) {
    k_omega<<<(H_env * W_env + 127) / 128, 128, 0, st>>>(H_env, W_env, omega);
    double *rowtot;
    cudaError_t e = cudaMalloc(&rowtot, sizeof(double) * H_env);
    if (e != cudaSuccess) return e;
    k_row_prefix1<<<H_env, 32, 0, st>>>(W_env, omega, omega_pref, rowtot);
    k_sum_serial<<<1, 32, 0, st>>>(H_env, rowtot, omega_total);
    e = cudaStreamSynchronize(st);
    cudaFree(rowtot);
    return e != cudaSuccess ? e : cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// per frame: channel sums of the uint8 input (irradiance mean add_attenuation.py:70, bg mean generator.py:462)
// ------------------------------------------------------------------------------------------
__global__ void k_stats(const uint8_t *bgr, int npix, unsigned long long *chan_sum) {
    int f = blockIdx.y;
    const uint8_t *p = bgr + (size_t)f * npix * 3;
    unsigned int s0 = 0, s1 = 0, s2 = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += gridDim.x * blockDim.x) {
        s0 += p[3 * i]; s1 += p[3 * i + 1]; s2 += p[3 * i + 2];
    }
    __shared__ unsigned int sh[3][8];
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { sh[0][w] = s0; sh[1][w] = s1; sh[2][w] = s2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        unsigned long long t = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) t += sh[threadIdx.x][k];
        atomicAdd(&chan_sum[f * 4 + threadIdx.x], t);      // integer: order independent, exact
    }
}

__global__ void k_stats_final(const unsigned long long *chan_sum, double *bg_sum, int F) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < F * 4) bg_sum[i] = (double)chan_sum[i] / 255.0;
}

// render_scale == 2: cv2.resize(bg / 255.0, (W, H)) with the default INTER_LINEAR switches to the exact
// 2x2 area average (resizeAreaFast_): sum = ((a + b) + c) + d in row-major cell order, times float 0.25
// (reference common/generator.py:352-355).  Output planar float64 plus per-block channel partial sums.
__global__ void __launch_bounds__(256) k_downscale2(const uint8_t *bgr, double *bgf, double *partial, int W, int H) {
    int f = blockIdx.y;
    const size_t np = (size_t)W * H;
    const uint8_t *src = bgr + (size_t)f * np * 4 * 3;
    double s0 = 0, s1 = 0, s2 = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < np; i += (size_t)gridDim.x * blockDim.x) {
        int y = (int)(i / W), x = (int)(i - (size_t)y * W);
        const uint8_t *p00 = src + ((size_t)(2 * y) * (2 * W) + 2 * x) * 3, *p10 = p00 + (size_t)(2 * W) * 3;
        double v[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            double a = rr_u8_unit(p00[c]), bq = rr_u8_unit(p00[3 + c]), cq = rr_u8_unit(p10[c]), d = rr_u8_unit(p10[3 + c]);
            double sum = ((a + bq) + cq) + d;
            v[c] = sum * 0.25f;
            bgf[((size_t)f * 3 + c) * np + i] = v[c];
        }
        s0 += v[0]; s1 += v[1]; s2 += v[2];
    }
    __shared__ double sh[3][8];
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2);
    int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { sh[0][w] = s0; sh[1][w] = s1; sh[2][w] = s2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0;
        for (int k = 0; k < 8; k++) t += sh[threadIdx.x][k];
        partial[((size_t)f * gridDim.x + blockIdx.x) * 4 + threadIdx.x] = t;
    }
}

__global__ void k_downscale2_final(const double *partial, double *bg_sum, int nblk) {
    int f = blockIdx.x;
    if (threadIdx.x < 3) {
        double t = 0;
        for (int k = 0; k < nblk; k++) t += partial[((size_t)f * nblk + k) * 4 + threadIdx.x];     // fixed order: deterministic
        bg_sum[f * 4 + threadIdx.x] = t;
    }
}

cudaError_t rr_launch_stats(const rr_frame_bufs &b, int F, int W, int H, int render_scale, double *bgf_out, cudaStream_t st) {
    if (render_scale == 2) {
        dim3 grid(64, F);
        // partial sums live at the head of the (not yet used) tile_sum buffer
        k_downscale2<<<grid, 256, 0, st>>>(b.bgr, bgf_out, b.tile_sum, W, H);
        k_downscale2_final<<<F, 32, 0, st>>>(b.tile_sum, b.bg_sum, 64);
        return cudaGetLastError();
    }
    cudaMemsetAsync(b.chan_sum, 0, sizeof(unsigned long long) * 4 * F, st);
    dim3 grid(64, F);
    k_stats<<<grid, 256, 0, st>>>(b.bgr, W * H, b.chan_sum);
    k_stats_final<<<(F * 4 + 127) / 128, 128, 0, st>>>(b.chan_sum, b.bg_sum, F);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// per frame: fog-like rain attenuation, one fused tile kernel  (add_attenuation.py:40-95)
//   f_ext  = float32 exp(-beta * depth/1000)                     (:43-48, float32 like numpy)
//   l_in_c = clip(beta_hg * E_c * (1 - f_ext), 0, 1)             (:53-72)
//   both blurred 25x25 sigma 25 REFLECT_101                       (:79-80)
//   l = clip(I * f_blur + l_in_blur, 0, 1)                        (:85-93)
// ------------------------------------------------------------------------------------------
#define FOG_TX 64
#ifndef FOG_TY
#define FOG_TY 32             // rows per tile, a multiple of 8: FOG_TX x (FOG_TY / 8) threads, 8 output rows per thread
#endif
#define FOG_THREADS (FOG_TX * (FOG_TY / 8))
#define FOG_MINB (FOG_THREADS <= 256 ? 2 : 1)
#define FOG_R 12
#define FOG_EW (FOG_TX + 2 * FOG_R)
#define FOG_EH (FOG_TY + 2 * FOG_R)
// Shared-memory layout.  In the row passes a thread produces 4 consecutive outputs from 28 consecutive inputs,
// so the lanes of a warp are 4 elements apart: unpadded, a float32 load hits 8 banks (4-way conflict) and a
// float64 load 4 bank pairs.  One padding element after every 4 (column x lives at x + x/4) makes the lane
// stride 5 elements, which is conflict free for both widths; the float32 row stride is 16 mod 32 words so
// that the two rows a warp covers use complementary banks.
#define FOG_PAD(x) ((x) + ((x) >> 2))
#define FOG_ES 112            // float32 row stride of E   (>= FOG_PAD(FOG_EW - 1) + 1, 16 mod 32)
#define FOG_FS 80             // row stride of FH / LH     (== FOG_PAD(FOG_TX), 16 mod 32)
#define FOG_LS 110            // float64 row stride of D   (>= FOG_PAD(FOG_EW - 1) + 1)
#define FOG_BYTES_A (sizeof(float) * FOG_EH * (FOG_ES + FOG_FS))     // E + FH, later LH (float64, FOG_EH x FOG_FS)
#define FOG_BYTES_B (sizeof(double) * FOG_EH * FOG_LS)              // D

// Sliding-window separable passes: each thread produces 4 consecutive outputs from 28 inputs held in
// registers; every output still accumulates its 25 products in the reference order.
// ROW: inputs are consecutive padded columns starting at a multiple of 4; else rows STRIDE apart.
// ROW: 0 = a column (elements STRIDE apart), 1 = a row in the padded layout, 2 = a row of the dense tile TMA delivered
// (16-byte aligned: seven 128-bit loads; lanes 4 elements apart then touch consecutive quads -- conflict free)
template <int ROW, int STRIDE>
__device__ __forceinline__ void fog_taps_f32(const float *in, double acc[4]) {
    // exact float32 products in float64, ascending taps.  The product of two float32 values is exact in
    // float64, so fma(k, v, a) == a + k * v bit for bit: one instruction per tap instead of two.
    double v[28];
    if (ROW == 2) {
#pragma unroll
        for (int q = 0; q < 7; q++) {
            const float4 t = ((const float4 *)in)[q];
            v[4 * q] = (double)t.x; v[4 * q + 1] = (double)t.y; v[4 * q + 2] = (double)t.z; v[4 * q + 3] = (double)t.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 28; i++) v[i] = (double)in[ROW ? FOG_PAD(i) : i * STRIDE];
    }
#pragma unroll
    for (int o = 0; o < 4; o++) {
        double a = c_k32d[0] * v[o];
#pragma unroll
        for (int t = 1; t < 25; t++) a = __fma_rn(c_k32d[t], v[o + t], a);
        acc[o] = a;
    }
}

// extinction f_ext = float32 exp(-beta * depth / 1000), once per pixel (add_attenuation.py:43-48); k_fog's
// overlapping tiles read it back (2.4 haloed reads per pixel) instead of re-evaluating the exponential
template <bool U16>
__global__ void __launch_bounds__(256) k_fext(const void *depth, float *fext, float neg_beta32, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // U16: the PNG's 16-bit samples; depth = sample.astype(float32) / 256 (generator.py:365) is exact in float32
    const float metres = U16 ? __fdiv_rn((float)((const uint16_t *)depth)[i], 256.f) : ((const float *)depth)[i];
    float d = __fdiv_rn(metres, 1000.f);                                    // add_attenuation.py:48 (float32)
    float xx = __fmul_rn(neg_beta32, d);
    fext[i] = (float)exp((double)xx);                                       // correctly rounded float32 exp ("canonical")
}

// per frame and channel: A_c = beta_hg * mean irradiance of the un-fogged image (add_attenuation.py:53,70) -- once, not
// once per thread of every tile (two float64 divisions each)
__global__ void k_fog_acs(const double *bg_sum, double *acs, rr_fog_consts fc, double npix, int F) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F * 4) return;
    double irr_mean = ((fc.irr_scale_num * bg_sum[i]) / fc.irr_den) / npix;
    acs[i] = fc.beta_hg * irr_mean;
}

// The TMA form of the fog stage reads the extinction from a REFLECT-PADDED plane: k_fext_pad materialises the
// BORDER_REFLECT_101 halo (FOG_R pixels on every side) once, with a row pitch that is a multiple of 16 bytes, so that
// every haloed tile of k_fog -- border tiles included -- is one box of a tensor map and no thread forms a reflected
// index.  Element (py, px) of the plane is f_ext(r101(py - FOG_R), r101(px - FOG_R)); pitch columns beyond W + 2 FOG_R
// are zero (tiles never consume them).  With uint16 depth the exponential is a 65536-entry table (k_fext_lut, built
// per camera): the correctly rounded float32 value of every possible sample.
__global__ void __launch_bounds__(256) k_fext_lut(float *lut, float neg_beta32) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 65536) return;
    const float metres = __fdiv_rn((float)i, 256.f);
    const float d = __fdiv_rn(metres, 1000.f);
    lut[i] = (float)exp((double)__fmul_rn(neg_beta32, d));
}

template <bool U16>
__global__ void __launch_bounds__(256) k_fext_pad(const void *depth, const float *lut, float *fextp, float neg_beta32, int W, int H, int Wp, int Hp) {
    const int f = blockIdx.z, py = blockIdx.y;
    const int px = blockIdx.x * blockDim.x + threadIdx.x;
    if (px >= Wp) return;
    // the source row is the block's (one reflected row index per block); columns reflect per thread
    const size_t row0 = ((size_t)f * H + r101(py - FOG_R, H)) * W;
    float v = 0.f;
    if (px < W + 2 * FOG_R) {
        const int sx = r101(px - FOG_R, W);
        if (U16) v = __ldg(&lut[__ldg((const uint16_t *)depth + row0 + sx)]);
        else {
            const float d = __fdiv_rn(((const float *)depth)[row0 + sx], 1000.f);
            v = (float)exp((double)__fmul_rn(neg_beta32, d));
        }
    }
    fextp[((size_t)f * Hp + py) * Wp + px] = v;
}

#define FOG_ED 88             // row stride of the dense extinction tile (== FOG_EW, 352 bytes)
#define FOG_BYTES_A_TMA (sizeof(float) * FOG_EH * (FOG_ED + FOG_FS))

template <bool TMA>
__global__ void __launch_bounds__(FOG_THREADS, FOG_MINB) k_fog(rr_frame_bufs b, rr_fog_consts fc, int W, int H, int skip_linear,
                                                               const __grid_constant__ CUtensorMap fmap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int ES = TMA ? FOG_ED : FOG_ES;
    constexpr size_t BYTES_A = TMA ? FOG_BYTES_A_TMA : FOG_BYTES_A;
    float *E = (float *)smem_raw;                         // [FOG_EH][ES]      extinction on the haloed tile (TMA: dense, as the box arrives)
    float *FH = E + FOG_EH * ES;                          // [FOG_EH][FOG_FS]  float32 row pass of f_ext
    double *LH = (double *)smem_raw;                      // [FOG_EH][FOG_FS]  float64 row pass; reuses E + FH once both are consumed
    double *D = (double *)(smem_raw + BYTES_A);           // [FOG_EH][FOG_LS]  1 - f_ext (float32 op, widened)
    uint8_t *IB = smem_raw + BYTES_A + FOG_BYTES_B;       // [FOG_TY][FOG_TX * 3]  the tile's image bytes
    __shared__ __align__(8) unsigned long long tile_bar;
    const int f = blockIdx.z;
    const int x0 = blockIdx.x * FOG_TX, y0 = blockIdx.y * FOG_TY;
    const int tid = threadIdx.x;
    const float *fext = b.fext + (size_t)f * W * H;          // tight plane (the non-TMA form)
    const uint8_t *bgr = b.bgr + (size_t)f * W * H * 3;
    double Acs[3];
#pragma unroll
    for (int c = 0; c < 3; c++) Acs[c] = b.acs[f * 4 + c];                      // beta_hg * E_c of the frame (k_fog_acs)
    // When beta_hg * E_c <= 1 for all channels the clip at :72 can only act on a negative 1 - f_ext (negative
    // depth), the three in-scatter images are the same image times a scalar, and one blur serves all three:
    // blur(A*d) = A*blur(d) up to float64 rounding (shortcut 3).  Otherwise each channel is
    // blurred on its own: l_in_c = clip(A_c * (1 - f_ext), 0, 1) is then formed while the row pass loads its inputs.
    const bool linear = Acs[0] >= 0 && Acs[0] <= 1 && Acs[1] >= 0 && Acs[1] <= 1 && Acs[2] >= 0 && Acs[2] <= 1;
    if (skip_linear && linear) return;                    // k_fog_roll renders this frame (block-uniform)
    // All global loads of the tile are issued up front, back to back (the kernel runs 4 warps per scheduler, too
    // few to hide a load that is consumed right away): the haloed extinction values, and the tile's own image
    // bytes, which wait in shared memory for the compose step at the very end.  A warp takes whole rows (the
    // reflected row index is warp-uniform, the reflected column indices are formed once per lane), so an element
    // costs a handful of instructions of index arithmetic.
    if (TMA) {
        // one elected thread asks the TMA unit for the haloed extinction tile (88 x 56 float32, one box of the padded
        // plane) while all threads fetch the tile's image bytes; then everybody waits on the mbarrier
        if (tid == 0) mbar_init(&tile_bar, 1);
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&tile_bar, (unsigned)(sizeof(float) * FOG_EH * FOG_ED));
            tma_load_3d(E, &fmap, x0, y0, b.frame0 + f, &tile_bar);
        }
        constexpr int NW = FOG_THREADS / 32, BR = FOG_TY / NW, BC = FOG_TX * 3 / 32;
        const int lane = tid & 31, warp = tid >> 5;
        if (!b.bgf) {
            uint8_t bv[BR][BC];
#pragma unroll
            for (int q = 0; q < BR; q++) {
                const int row = warp + NW * q;
                const bool rok = y0 + row < H;
                const uint8_t *src = bgr + ((size_t)(rok ? y0 + row : 0) * W + x0) * 3;
#pragma unroll
                for (int j = 0; j < BC; j++) {
                    const int col = lane + 32 * j;
                    bv[q][j] = (rok && x0 * 3 + col < W * 3) ? src[col] : (uint8_t)0;
                }
            }
#pragma unroll
            for (int q = 0; q < BR; q++)
#pragma unroll
                for (int j = 0; j < BC; j++) IB[(warp + NW * q) * (FOG_TX * 3) + lane + 32 * j] = bv[q][j];
        }
        mbar_wait(&tile_bar, 0);
        // 1 - f_ext of the haloed tile, widened, into the padded float64 layout of the row pass
        for (int i = tid; i < FOG_EH * FOG_EW; i += FOG_THREADS) {
            const int ey = i / FOG_EW, ex = i - ey * FOG_EW;
            double d = (double)(1.0f - E[ey * FOG_ED + ex]);                // (1 - f_ext) is a float32 op in numpy (:71)
            if (linear) d = d < 0 ? 0 : (d > 1 ? 1 : d);                    // clip(A d, 0, 1) = A clip(d, 0, 1) for 0 <= A <= 1
            D[ey * FOG_LS + FOG_PAD(ex)] = d;
        }
    } else {
        constexpr int NW = FOG_THREADS / 32;                                // warps
        constexpr int ER = (FOG_EH + NW - 1) / NW;                          // extinction rows per warp
        constexpr int EC = (FOG_EW + 31) / 32;                              // column steps per row
        constexpr int BR = FOG_TY / NW;                                     // image rows per warp
        constexpr int BC = FOG_TX * 3 / 32;                                 // byte steps per image row
        const int lane = tid & 31, warp = tid >> 5;
        int gxs[EC], pex[EC];
#pragma unroll
        for (int j = 0; j < EC; j++) {
            const int ex = lane + 32 * j;
            const int gx = r101(x0 + ex - FOG_R, W);
            gxs[j] = (ex < FOG_EW && gx >= 0 && gx < W) ? gx : -1;
            pex[j] = ex < FOG_EW ? FOG_PAD(ex) : -1;
        }
        float ev[ER][EC];
        uint8_t bv[BR][BC];
#pragma unroll
        for (int q = 0; q < ER; q++) {
            const int ey = warp + NW * q;
            const int gy = r101(y0 + ey - FOG_R, H);
            const bool rok = ey < FOG_EH && gy >= 0 && gy < H;
            const float *src = fext + (size_t)(rok ? gy : 0) * W;
#pragma unroll
            for (int j = 0; j < EC; j++) ev[q][j] = (rok && gxs[j] >= 0) ? src[gxs[j]] : 0.f;                       // k_fext
        }
        if (!b.bgf) {
#pragma unroll
            for (int q = 0; q < BR; q++) {
                const int row = warp + NW * q;
                const bool rok = y0 + row < H;
                const uint8_t *src = bgr + ((size_t)(rok ? y0 + row : 0) * W + x0) * 3;
#pragma unroll
                for (int j = 0; j < BC; j++) {
                    const int col = lane + 32 * j;
                    bv[q][j] = (rok && x0 * 3 + col < W * 3) ? src[col] : (uint8_t)0;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < ER; q++) {
            const int ey = warp + NW * q;
            if (ey < FOG_EH) {
#pragma unroll
                for (int j = 0; j < EC; j++) {
                    if (pex[j] >= 0) {
                        const float v = ev[q][j];
                        E[ey * FOG_ES + pex[j]] = v;
                        double d = (double)(1.0f - v);                      // (1 - f_ext) is a float32 op in numpy (:71)
                        if (linear) d = d < 0 ? 0 : (d > 1 ? 1 : d);        // clip(A d, 0, 1) = A clip(d, 0, 1) for 0 <= A <= 1
                        D[ey * FOG_LS + pex[j]] = d;
                    }
                }
            }
        }
        if (!b.bgf) {
#pragma unroll
            for (int q = 0; q < BR; q++)
#pragma unroll
                for (int j = 0; j < BC; j++) IB[(warp + NW * q) * (FOG_TX * 3) + lane + 32 * j] = bv[q][j];
        }
    }
    __syncthreads();
    // float32 row pass of f_ext: 4 outputs per task
    for (int i = tid; i < FOG_EH * (FOG_TX / 4); i += FOG_THREADS) {
        int ey = i / (FOG_TX / 4), ox = (i - ey * (FOG_TX / 4)) * 4;
        double a[4];
        if (TMA) fog_taps_f32<2, 0>(E + ey * FOG_ED + ox, a);
        else fog_taps_f32<1, 0>(E + ey * FOG_ES + FOG_PAD(ox), a);
#pragma unroll
        for (int o = 0; o < 4; o++) FH[ey * FOG_FS + FOG_PAD(ox) + o] = (float)a[o];
    }
    __syncthreads();
    // float32 column pass: thread = (column, block of 8 rows)
    const int cx = tid & (FOG_TX - 1), cy0 = (tid / FOG_TX) * 8;
    const int pcx = FOG_PAD(cx);
    float fb[8];
    {
        double a[4];
        fog_taps_f32<0, FOG_FS>(FH + cy0 * FOG_FS + pcx, a);
#pragma unroll
        for (int o = 0; o < 4; o++) fb[o] = (float)a[o];
        fog_taps_f32<0, FOG_FS>(FH + (cy0 + 4) * FOG_FS + pcx, a);
#pragma unroll
        for (int o = 0; o < 4; o++) fb[4 + o] = (float)a[o];
    }
    if (b.fblur) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            int gy = y0 + cy0 + k, gx = x0 + cx;
            if (gy < H && gx < W) b.fblur[(size_t)f * W * H + (size_t)gy * W + gx] = fb[k];
        }
    }
    const int npass = linear ? 1 : 3;
    for (int c = 0; c < npass; c++) {
        const double Ac = Acs[c];
        __syncthreads();            // E and FH consumed (first pass) / LH of the previous channel consumed
        // float64 row pass, cv::RowFilter order: s = k[0]*x[0]; s += k[t]*x[t] ... -- OpenCV's AVX2-dispatched build of that
        // scalar loop is contracted to fused multiply-adds by its compiler (measured: the FMA chain reproduces
        // cv2.sepFilter2D bit for bit on float64, the separate multiply-add does not), so the chain is fused here too.
        // The column filter (SymmColumnFilter) is not contracted in that build and stays multiply + add.
        for (int i = tid; i < FOG_EH * (FOG_TX / 4); i += FOG_THREADS) {
            int ey = i / (FOG_TX / 4), ox = (i - ey * (FOG_TX / 4)) * 4;
            const double *row = D + ey * FOG_LS + FOG_PAD(ox);
            double v[28];
#pragma unroll
            for (int k = 0; k < 28; k++) v[k] = row[FOG_PAD(k)];
            if (!linear) {
#pragma unroll
                for (int k = 0; k < 28; k++) { double li = Ac * v[k]; v[k] = li < 0 ? 0 : (li > 1 ? 1 : li); }   // :71-72
            }
#pragma unroll
            for (int o = 0; o < 4; o++) {
                double a = c_k64[0] * v[o];
#pragma unroll
                for (int t = 1; t < 25; t++) a = __fma_rn(c_k64[t], v[o + t], a);
                LH[ey * FOG_FS + FOG_PAD(ox) + o] = a;
            }
        }
        __syncthreads();
        // float64 column pass, cv::SymmColumnFilter order: k[c]*x[0] + sum_t k[c+t]*(x[+t] + x[-t]); then compose
        {
            double v[32];
#pragma unroll
            for (int k = 0; k < 32; k++) v[k] = LH[(cy0 + k) * FOG_FS + pcx];
#pragma unroll
            for (int o = 0; o < 8; o++) {
                double acc = c_k64[12] * v[o + 12];
#pragma unroll
                for (int t = 1; t <= 12; t++) acc += c_k64[12 + t] * (v[o + 12 + t] + v[o + 12 - t]);
                int gy = y0 + cy0 + o, gx = x0 + cx;
                if (gy < H && gx < W) {
                    size_t pix = (size_t)gy * W + gx;
                    unsigned packed = 0;
                    for (int cc = linear ? 0 : c; cc < (linear ? 3 : c + 1); cc++) {
                        double I = b.bgf ? b.bgf[((size_t)f * 3 + cc) * W * H + pix] : rr_u8_unit(IB[((cy0 + o) * FOG_TX + cx) * 3 + cc]);   // generator.py:352-355
                        double lin_in = linear ? Acs[cc] * acc : acc;
                        double l = I * (double)fb[o] + lin_in;               // :85
                        l = l < 0 ? 0 : (l > 1 ? 1 : l);
                        b.rainy[((size_t)f * 3 + cc) * W * H + pix] = l;
                        const unsigned q = (unsigned)(uint8_t)(l * 255);     // bad_weather.py:744
                        if (linear) packed |= q << (8 * cc);
                        else b.bg8[((size_t)f * W * H + pix) * 4 + cc] = (uint8_t)q;
                    }
                    if (linear) ((unsigned *)b.bg8)[(size_t)f * W * H + pix] = packed;
                }
            }
        }
    }
}

// ---- rolling form of the fog stage --------------------------------------------------------------------------------------
// k_fog spends 1.75 row passes per output row: every 64 x 32 tile recomputes the row passes of its 24 halo rows, which its
// vertical neighbours compute too.  Here a CTA owns a 64-column strip and walks DOWN it: the row-pass results (float32 pass of
// f_ext, float64 pass of 1 - f_ext) live in two 32-row slots of a ring in shared memory, every step row-filters ONE new block
// of 32 padded rows (one TMA box of the reflect-padded plane, 88 x 32) and column-filters the 56-row window that the ring now
// holds -- one row pass per output row, a quarter of the stage's float64 work gone.  A strip is cut into FOGR_SEG-tile
// segments so that the grid keeps the machine full (each segment pays one extra block).  Only the frames whose in-scatter
// images are one image times a scalar ("linear", the common case) take this form; the others leave at once and are
// rendered by k_fog, which in turn skips the linear frames.
#ifndef FOGR_SEG
#define FOGR_SEG 6            // tiles per CTA (H = 375: 12 tiles -> 2 segments)
#endif
#define FOGR_BR 32            // rows per block = FOG_TY
#define FOGR_BYTES_FH (sizeof(float) * 2 * FOGR_BR * FOG_FS)          // float32 row-pass ring   [2][32][FOG_FS]
#define FOGR_BYTES_LH (sizeof(double) * 2 * FOGR_BR * FOG_FS)         // float64 row-pass ring   [2][32][FOG_FS]
#define FOGR_BYTES_E (sizeof(float) * FOGR_BR * FOG_ED)               // the block as TMA delivers it [32][88]
#define FOGR_BYTES_D (sizeof(double) * FOGR_BR * FOG_LS)              // 1 - f_ext of the block, widened, padded layout
#define FOGR_SMEM (FOGR_BYTES_E + FOGR_BYTES_FH + FOGR_BYTES_LH + FOGR_BYTES_D + FOG_TY * FOG_TX * 3)

__device__ __forceinline__ bool fog_frame_linear(const rr_frame_bufs &b, int f, double Acs[3]) {
#pragma unroll
    for (int c = 0; c < 3; c++) Acs[c] = b.acs[f * 4 + c];
    return Acs[0] >= 0 && Acs[0] <= 1 && Acs[1] >= 0 && Acs[1] <= 1 && Acs[2] >= 0 && Acs[2] <= 1;
}

__global__ void __launch_bounds__(FOG_THREADS, FOG_MINB) k_fog_roll(rr_frame_bufs b, rr_fog_consts fc, int W, int H, int tiles_y,
                                                                    const __grid_constant__ CUtensorMap fmap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *E = (float *)smem_raw;                                            // [32][FOG_ED]
    float *FH = (float *)(smem_raw + FOGR_BYTES_E);                          // [2][32][FOG_FS]
    double *LH = (double *)(smem_raw + FOGR_BYTES_E + FOGR_BYTES_FH);        // [2][32][FOG_FS]
    double *D = (double *)(smem_raw + FOGR_BYTES_E + FOGR_BYTES_FH + FOGR_BYTES_LH);   // [32][FOG_LS]
    uint8_t *IB = smem_raw + FOGR_BYTES_E + FOGR_BYTES_FH + FOGR_BYTES_LH + FOGR_BYTES_D;
    __shared__ __align__(8) unsigned long long blk_bar;
    const int f = blockIdx.z, tid = threadIdx.x;
    double Acs[3];
    if (!fog_frame_linear(b, f, Acs)) return;                                // k_fog renders this frame (block-uniform)
    const int x0 = blockIdx.x * FOG_TX;
    const int t_first = blockIdx.y * FOGR_SEG, t_last = (t_first + FOGR_SEG < tiles_y ? t_first + FOGR_SEG : tiles_y) - 1;
    const uint8_t *bgr = b.bgr + (size_t)f * W * H * 3;
    constexpr int NW = FOG_THREADS / 32, BR = FOG_TY / NW, BC = FOG_TX * 3 / 32;
    const int lane = tid & 31, warp = tid >> 5;
    const int cx = tid & (FOG_TX - 1), cy0 = (tid / FOG_TX) * 8, pcx = FOG_PAD(cx);
    if (tid == 0) mbar_init(&blk_bar, 1);
    __syncthreads();
    unsigned uses = 0;                                                       // completed waits on blk_bar (its phase parity)
    auto request = [&](int blk) {                                            // one thread: padded rows [32 blk, 32 blk + 32) of the strip
        mbar_expect_tx(&blk_bar, (unsigned)FOGR_BYTES_E);
        tma_load_3d(E, &fmap, x0, blk * FOGR_BR, b.frame0 + f, &blk_bar);
    };
    // row passes of the block that sits in E, into ring slot `slot`
    auto row_passes = [&](int slot) {
        mbar_wait(&blk_bar, uses & 1u);
        uses++;
        for (int i = tid; i < FOGR_BR * FOG_EW; i += FOG_THREADS) {
            const int ey = i / FOG_EW, ex = i - ey * FOG_EW;
            double d = (double)(1.0f - E[ey * FOG_ED + ex]);                 // (1 - f_ext) is a float32 op in numpy (:71)
            d = d < 0 ? 0 : (d > 1 ? 1 : d);                                 // clip(A d, 0, 1) = A clip(d, 0, 1) for 0 <= A <= 1
            D[ey * FOG_LS + FOG_PAD(ex)] = d;
        }
        float *fh = FH + slot * FOGR_BR * FOG_FS;
        for (int i = tid; i < FOGR_BR * (FOG_TX / 4); i += FOG_THREADS) {
            const int ey = i / (FOG_TX / 4), ox = (i - ey * (FOG_TX / 4)) * 4;
            double a[4];
            fog_taps_f32<2, 0>(E + ey * FOG_ED + ox, a);
#pragma unroll
            for (int o = 0; o < 4; o++) fh[ey * FOG_FS + FOG_PAD(ox) + o] = (float)a[o];
        }
        __syncthreads();                                                     // E consumed, D complete
    };
    auto row_pass64 = [&](int slot) {
        double *lh = LH + slot * FOGR_BR * FOG_FS;
        for (int i = tid; i < FOGR_BR * (FOG_TX / 4); i += FOG_THREADS) {
            const int ey = i / (FOG_TX / 4), ox = (i - ey * (FOG_TX / 4)) * 4;
            const double *row = D + ey * FOG_LS + FOG_PAD(ox);
            double v[28];
#pragma unroll
            for (int k = 0; k < 28; k++) v[k] = row[FOG_PAD(k)];
#pragma unroll
            for (int o = 0; o < 4; o++) {
                double a = c_k64[0] * v[o];
#pragma unroll
                for (int t = 1; t < 25; t++) a = __fma_rn(c_k64[t], v[o + t], a);   // cv::RowFilter, fused like OpenCV's build (see k_fog)
                lh[ey * FOG_FS + FOG_PAD(ox) + o] = a;
            }
        }
    };
    // prologue: block t_first into slot (t_first & 1)
    if (tid == 0) request(t_first);
    row_passes(t_first & 1);
    if (tid == 0) request(t_first + 1);                                      // E is free again: the next block travels under the float64 pass
    row_pass64(t_first & 1);
    __syncthreads();                                                         // D consumed: the loop's first row_passes overwrites it (racecheck)
    for (int t = t_first; t <= t_last; t++) {
        const int y0 = t * FOG_TY;
        // the tile's image bytes, requested now and parked in shared memory until the compose step
        uint8_t bv[BR][BC];
        if (!b.bgf) {
#pragma unroll
            for (int q = 0; q < BR; q++) {
                const int row = warp + NW * q;
                const bool rok = y0 + row < H;
                const uint8_t *src = bgr + ((size_t)(rok ? y0 + row : 0) * W + x0) * 3;
#pragma unroll
                for (int j = 0; j < BC; j++) {
                    const int col = lane + 32 * j;
                    bv[q][j] = (rok && x0 * 3 + col < W * 3) ? src[col] : (uint8_t)0;
                }
            }
        }
        // block t + 1 (already requested) into the other slot
        const int s1 = (t + 1) & 1;
        row_passes(s1);
        if (t < t_last && tid == 0) request(t + 2);
        row_pass64(s1);
        if (!b.bgf) {
#pragma unroll
            for (int q = 0; q < BR; q++)
#pragma unroll
                for (int j = 0; j < BC; j++) IB[(warp + NW * q) * (FOG_TX * 3) + lane + 32 * j] = bv[q][j];
        }
        __syncthreads();                                                     // both slots and IB complete
        // window row r (0 .. 55) of tile t: block t + (r >> 5), row r & 31 of its slot
        const int tpar = t & 1;
        float fb[8];
        {
            double v[32];
#pragma unroll
            for (int k = 0; k < 32; k++) {
                const int r = cy0 + k;
                v[k] = (double)FH[(((tpar + (r >> 5)) & 1) * FOGR_BR + (r & 31)) * FOG_FS + pcx];
            }
#pragma unroll
            for (int o = 0; o < 8; o++) {
                double a = c_k32d[0] * v[o];
#pragma unroll
                for (int tt = 1; tt < 25; tt++) a = __fma_rn(c_k32d[tt], v[o + tt], a);   // exact float32 products, ascending taps (fog_taps_f32)
                fb[o] = (float)a;
            }
        }
        if (b.fblur) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int gy = y0 + cy0 + k, gx = x0 + cx;
                if (gy < H && gx < W) b.fblur[(size_t)f * W * H + (size_t)gy * W + gx] = fb[k];
            }
        }
        {
            double v[32];
#pragma unroll
            for (int k = 0; k < 32; k++) {
                const int r = cy0 + k;
                v[k] = LH[(((tpar + (r >> 5)) & 1) * FOGR_BR + (r & 31)) * FOG_FS + pcx];
            }
#pragma unroll
            for (int o = 0; o < 8; o++) {
                double acc = c_k64[12] * v[o + 12];                          // cv::SymmColumnFilter order
#pragma unroll
                for (int tt = 1; tt <= 12; tt++) acc += c_k64[12 + tt] * (v[o + 12 + tt] + v[o + 12 - tt]);
                const int gy = y0 + cy0 + o, gx = x0 + cx;
                if (gy < H && gx < W) {
                    const size_t pix = (size_t)gy * W + gx;
                    unsigned packed = 0;
#pragma unroll
                    for (int cc = 0; cc < 3; cc++) {
                        const double I = b.bgf ? b.bgf[((size_t)f * 3 + cc) * W * H + pix] : rr_u8_unit(IB[((cy0 + o) * FOG_TX + cx) * 3 + cc]);
                        double l = I * (double)fb[o] + Acs[cc] * acc;        // add_attenuation.py:85
                        l = l < 0 ? 0 : (l > 1 ? 1 : l);
                        b.rainy[((size_t)f * 3 + cc) * W * H + pix] = l;
                        packed |= (unsigned)(uint8_t)(l * 255) << (8 * cc);  // bad_weather.py:744
                    }
                    ((unsigned *)b.bg8)[(size_t)f * W * H + pix] = packed;
                }
            }
        }
        __syncthreads();                                                     // the window is consumed: slot (t & 1) and IB may be overwritten
    }
}

cudaError_t rr_launch_fext_lut(float *lut, float neg_beta32, cudaStream_t st) {
    k_fext_lut<<<256, 256, 0, st>>>(lut, neg_beta32);
    return cudaGetLastError();
}

// The extinction plane of the TMA form needs only the depth: it may run on another stream beside k_stats (rr_launch_fog is
// then told that the plane is already on its way: fext_done).
cudaError_t rr_launch_fext_pad(const rr_frame_bufs &b, const rr_fog_consts &fc, int F, int W, int H, cudaStream_t st) {
    dim3 g((b.fext_Wp + 255) / 256, b.fext_Hp, F);
    float *dst = b.fext + (size_t)b.frame0 * b.fext_Hp * b.fext_Wp;     // the tensor map spans the whole plane stack: frame0 + f
    if (b.depth_u16) k_fext_pad<true><<<g, 256, 0, st>>>(b.depth, b.fext_lut, dst, fc.neg_beta32, W, H, b.fext_Wp, b.fext_Hp);
    else k_fext_pad<false><<<g, 256, 0, st>>>(b.depth, b.fext_lut, dst, fc.neg_beta32, W, H, b.fext_Wp, b.fext_Hp);
    return cudaGetLastError();
}

cudaError_t rr_launch_fog(const rr_frame_bufs &b, const rr_fog_consts &fc, int F, int W, int H, const CUtensorMap *fmap, const CUtensorMap *fmap_roll,
                          bool fext_done, cudaStream_t st) {
    static_assert(FOG_ES >= FOG_PAD(FOG_EW - 1) + 1 && FOG_LS >= FOG_PAD(FOG_EW - 1) + 1 && FOG_FS >= FOG_PAD(FOG_TX - 1) + 1, "fog strides");
    static_assert(sizeof(double) * FOG_EH * FOG_FS <= FOG_BYTES_A && sizeof(double) * FOG_EH * FOG_FS <= FOG_BYTES_A_TMA, "LH must fit in the E + FH region");
    static_assert(FOG_ED == FOG_EW && (FOG_ED * sizeof(float)) % 16 == 0 && FOG_BYTES_A_TMA % 8 == 0, "dense tile layout");
    k_fog_acs<<<(F * 4 + 127) / 128, 128, 0, st>>>(b.bg_sum, b.acs, fc, (double)W * (double)H, F);
    dim3 grid((W + FOG_TX - 1) / FOG_TX, (H + FOG_TY - 1) / FOG_TY, F);
    if (fmap) {
        if (!fext_done) {
            cudaError_t e = rr_launch_fext_pad(b, fc, F, W, H, st);
            if (e != cudaSuccess) return e;
        }
        const size_t smem = FOG_BYTES_A_TMA + FOG_BYTES_B + FOG_TY * FOG_TX * 3;
        if (b.fog_roll) {
            // linear frames by the rolling kernel, the others by k_fog: each leaves the other's frames at once
            dim3 gr(grid.x, (grid.y + FOGR_SEG - 1) / FOGR_SEG, F);
            k_fog_roll<<<gr, FOG_THREADS, FOGR_SMEM, st>>>(b, fc, W, H, (int)grid.y, *fmap_roll);
        }
        k_fog<true><<<grid, FOG_THREADS, smem, st>>>(b, fc, W, H, b.fog_roll, *fmap);
    } else {
        const size_t n = (size_t)F * W * H;
        float *dst = b.fext + (size_t)b.frame0 * W * H;
        rr_frame_bufs v = b;
        v.fext = dst;
        if (b.depth_u16) k_fext<true><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(b.depth, dst, fc.neg_beta32, n);
        else k_fext<false><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(b.depth, dst, fc.neg_beta32, n);
        const size_t smem = FOG_BYTES_A + FOG_BYTES_B + FOG_TY * FOG_TX * 3;
        CUtensorMap dummy;
        memset(&dummy, 0, sizeof(dummy));
        k_fog<false><<<grid, FOG_THREADS, smem, st>>>(v, fc, W, H, 0, dummy);
    }
    return cudaGetLastError();
}

__global__ void k_planar_to_bg8(const double *planar, uint8_t *bg8, int npix, int F) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)F * npix) return;
    int f = (int)(i / npix);
    size_t pix = i - (size_t)f * npix;
    unsigned packed = 0;
    for (int c = 0; c < 3; c++) packed |= (unsigned)(uint8_t)(planar[((size_t)f * 3 + c) * npix + pix] * 255) << (8 * c);
    ((unsigned *)bg8)[i] = packed;
}

cudaError_t rr_launch_planar_to_bg8(const double *planar, uint8_t *bg8, int F, int W, int H, cudaStream_t st) {
    size_t n = (size_t)F * W * H;
    k_planar_to_bg8<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(planar, bg8, W * H, F);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// per frame: environment map = gather + 15x15 uint8 blur on the hole pixels  (bad_weather.py:742-819)
// ------------------------------------------------------------------------------------------
#define ENV_TX 64
#define ENV_TY 16
#define ENV_R 7
// Gather through the static source-index table and, for the never-written pixels, OpenCV's fixed-point
// 15x15 Gaussian of the gathered map (bad_weather.py:814-817), in one pass: tiles without holes are a pure gather.
// Pixels travel as one 32-bit word (B, G, R, 0): one load and one store per pixel instead of three each.
// per-camera: does the tile contain a never-written pixel?  (the answer is the same for every frame)
__global__ void __launch_bounds__(256) k_env_tile_flags(const uint8_t *env_written, uint8_t *tile_hole, int H, int W_env) {
    __shared__ int any_hole;
    const int x0 = blockIdx.x * ENV_TX, y0 = blockIdx.y * ENV_TY;
    if (threadIdx.x == 0) any_hole = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < ENV_TX * ENV_TY; i += 256) {
        int oy = i / ENV_TX, ox = i - oy * ENV_TX;
        int gy = y0 + oy, gx = x0 + ox;
        if (gy < H && gx < W_env && !env_written[(size_t)gy * W_env + gx]) any_hole = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) tile_hole[blockIdx.y * gridDim.x + blockIdx.x] = any_hole ? 1 : 0;
}

cudaError_t rr_launch_env_tile_flags(const uint8_t *env_written, uint8_t *tile_hole, int H, int W_env, cudaStream_t st) {
    dim3 g((W_env + ENV_TX - 1) / ENV_TX, (H + ENV_TY - 1) / ENV_TY);
    k_env_tile_flags<<<g, 256, 0, st>>>(env_written, tile_hole, H, W_env);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) k_env_map(const uint8_t *bg8, const int32_t *env_src, const uint8_t *env_written,
                                                 const uint8_t *tile_hole, uint8_t *env8, int H, int W_env, int pitch, int npix_img) {
    __shared__ unsigned in[ENV_TY + 2 * ENV_R][ENV_TX + 2 * ENV_R];
    __shared__ unsigned short hp[ENV_TY + 2 * ENV_R][ENV_TX][3];
    const int f = blockIdx.z;
    const int x0 = blockIdx.x * ENV_TX, y0 = blockIdx.y * ENV_TY;
    const unsigned *img = (const unsigned *)bg8 + (size_t)f * npix_img;
    unsigned *out = (unsigned *)env8 + (size_t)f * H * pitch;      // rows of `pitch` pixels (16-byte aligned for the bulk copies of k_env_prefix)
    const bool any_hole = tile_hole[blockIdx.y * gridDim.x + blockIdx.x] != 0;      // block-uniform, static per camera
    if (any_hole) {
        for (int i = threadIdx.x; i < (ENV_TY + 2 * ENV_R) * (ENV_TX + 2 * ENV_R); i += 256) {
            int ey = i / (ENV_TX + 2 * ENV_R), ex = i - ey * (ENV_TX + 2 * ENV_R);
            int gy = r101(y0 + ey - ENV_R, H), gx = r101(x0 + ex - ENV_R, W_env);
            unsigned v = 0;
            if (gy >= 0 && gy < H && gx >= 0 && gx < W_env) {
                int sidx = env_src[(size_t)gy * W_env + gx];
                if (sidx >= 0) v = img[sidx];
            }
            in[ey][ex] = v;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < (ENV_TY + 2 * ENV_R) * ENV_TX; i += 256) {
            int ey = i / ENV_TX, ox = i - ey * ENV_TX;
            unsigned s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
            for (int t = 0; t < 15; t++) {
                const unsigned v = in[ey][ox + t];
                s0 += c_k15[t] * (v & 255u); s1 += c_k15[t] * ((v >> 8) & 255u); s2 += c_k15[t] * ((v >> 16) & 255u);
            }
            hp[ey][ox][0] = (unsigned short)s0; hp[ey][ox][1] = (unsigned short)s1; hp[ey][ox][2] = (unsigned short)s2;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < ENV_TX * ENV_TY; i += 256) {
        int oy = i / ENV_TX, ox = i - oy * ENV_TX;
        int gy = y0 + oy, gx = x0 + ox;
        if (gy >= H || gx >= W_env) continue;
        const size_t pix = (size_t)gy * W_env + gx, opix = (size_t)gy * pitch + gx;
        if (env_written[pix]) {
            if (any_hole) out[opix] = in[oy + ENV_R][ox + ENV_R];
            else {
                int sidx = env_src[pix];
                out[opix] = sidx >= 0 ? img[sidx] : 0u;
            }
        } else {
            unsigned s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
            for (int t = 0; t < 15; t++) {
                s0 += c_k15[t] * hp[oy + t][ox][0]; s1 += c_k15[t] * hp[oy + t][ox][1]; s2 += c_k15[t] * hp[oy + t][ox][2];
            }
            out[opix] = ((s0 + 32768u) >> 16) | (((s1 + 32768u) >> 16) << 8) | (((s2 + 32768u) >> 16) << 16);
        }
    }
}

// one block per (row, frame): xyY, solid-angle weighting, row prefix sums  (generator.py:407-408,
// bad_weather.py:393-395).  pref is interleaved: [F][H][W_env+1][4] = prefix of (w*x, w*y, w*Y, w), so one
// 32-byte sector serves a span end point in k_setup.
//
// The row is cut into tiles of EP_TILE pixels.  Every thread owns EP_PER CONSECUTIVE pixels: it converts them
// (division-free xyY, rr_cvmath.h), keeps the running sums in registers (a serial prefix costs one add per
// value instead of a five-step shuffle scan) and parks its local exclusive prefixes in shared memory; one
// shuffle scan over the thread totals and the warp totals gives every thread its offset; the tile then
// leaves shared memory as fully coalesced 16-byte stores.  The row's pixels (32-bit words) arrive the same way
// (16-byte coalesced loads into shared memory, then each thread reads its 8 words).  Fixed tree: deterministic.
#define EP_THREADS 128
#define EP_PER 8
#define EP_TILE (EP_THREADS * EP_PER)
#define EP_WARPS (EP_THREADS / 32)
// shared staging of the tile's prefixes: 32 bytes per pixel plus 16 bytes after every 8 pixels, so that the
// per-thread 16-byte stores (stride 272 bytes between lanes) and the per-warp linear reads are conflict free
#ifndef RR_PREF_N
#define RR_PREF_N 4           // doubles per prefix entry: 4 = (w x, w y, w Y, w): one 32-byte sector per span end point; 3 drops w (k_setup then reads the
                              // per-camera prefix of w): measured on B200, -0.03 ms in k_env_prefix, +0.05 ms in k_setup -- no gain, 4 stays
#endif
#define EP_ENTRY (RR_PREF_N * 8)
#define EP_STAGE_BYTES (EP_TILE * EP_ENTRY + EP_THREADS * 16)
// BULK: the row's pixels arrive by 1-D bulk copies (cp.async.bulk, the TMA unit) into two alternating shared-memory buffers,
// signalled on mbarriers: tile t + 1 is in flight while tile t is converted.  Needs 16-byte aligned rows: env8 carries a
// pitch of a multiple of 4 pixels.  Otherwise the register-staged form (global -> registers -> shared, next tile prefetched).
template <bool BULK>
__global__ void __launch_bounds__(EP_THREADS) k_env_prefix(const uint8_t *env8, const double *omega, double *pref, double *rowtot,
                                                           int H, int W_env, int pitch) {
    __shared__ __align__(128) unsigned char s_bytes[BULK ? 2 * EP_TILE * 4 : EP_TILE * 4 + 32];
    __shared__ __align__(8) unsigned long long s_bar[2];
    __shared__ __align__(16) unsigned char s_stage[EP_STAGE_BYTES];
    __shared__ __align__(16) double s_toff[EP_THREADS][4];
    __shared__ double s_wtot[EP_WARPS][4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = blockIdx.x, f = blockIdx.y;
    const uint8_t *row = env8 + ((size_t)f * H + r) * pitch * 4;          // pixels are 32-bit words (B, G, R, 0)
    const double *om = omega + (size_t)r * W_env;
    double *prow = pref + ((size_t)f * H + r) * (W_env + 1) * RR_PREF_N;
    double cx = 0, cy = 0, cY = 0, cw = 0;          // carry: prefix of everything left of the tile
    // The pixels of a tile travel global -> registers -> shared memory, and the registers are refilled with the NEXT
    // tile's pixels right after they have been parked, so that load runs under this tile's arithmetic.
    constexpr int EP_VEC = (EP_TILE * 4 + 16 + 15) / 16 / EP_THREADS + 1;      // 16-byte vectors per thread (3)
    uint4 pv[EP_VEC];
    auto tile_src = [&](int c0, int *shift, int *nvec) -> const uint4 * {
        const int n = (W_env - c0) < EP_TILE ? (W_env - c0) : EP_TILE;
        const uint8_t *gsrc = row + (size_t)c0 * 4;
        *shift = (int)((size_t)gsrc & 15);                                 // a multiple of 4
        *nvec = (*shift + n * 4 + 15) >> 4;
        return (const uint4 *)(gsrc - *shift);                             // the buffers carry 256 bytes of slack
    };
    auto issue_bulk = [&](int t) {                  // one thread: tile t -> buffer t & 1
        const int c0 = t * EP_TILE;
        const unsigned bytes = (unsigned)(((pitch - c0) < EP_TILE ? (pitch - c0) : EP_TILE) * 4);
        mbar_expect_tx(&s_bar[t & 1], bytes);
        bulk_load_1d(s_bytes + (size_t)(t & 1) * EP_TILE * 4, row + (size_t)c0 * 4, bytes, &s_bar[t & 1]);
    };
    if (BULK) {
        if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); }
        __syncthreads();
        if (tid == 0) issue_bulk(0);
    } else {
        int sh0, nv0;
        const uint4 *g0 = tile_src(0, &sh0, &nv0);
#pragma unroll
        for (int q = 0; q < EP_VEC; q++) { const int i = tid + q * EP_THREADS; pv[q] = i < nv0 ? g0[i] : make_uint4(0, 0, 0, 0); }
    }
    int tile = 0;
    for (int c0 = 0; c0 < W_env; c0 += EP_TILE, tile++) {
        const int n = (W_env - c0) < EP_TILE ? (W_env - c0) : EP_TILE;
        int shift = 0, nvec = 0;
        if (!BULK) tile_src(c0, &shift, &nvec);
        // the thread's solid angles are requested now and consumed after the staging barrier
        const int px0 = tid * EP_PER;
        double wv[EP_PER];
#pragma unroll
        for (int k = 0; k < EP_PER; k++) wv[k] = (px0 + k < n) ? om[c0 + px0 + k] : 0.0;
        __syncthreads();                            // previous tile fully written out (s_stage, s_toff, s_bytes)
        if (BULK) {
            // the buffer of tile - 1 is free (barrier above): send tile + 1 into it, then wait for this tile's bytes
            if (tid == 0 && c0 + EP_TILE < W_env) issue_bulk(tile + 1);
            mbar_wait(&s_bar[tile & 1], (unsigned)((tile >> 1) & 1));
        } else {
#pragma unroll
            for (int q = 0; q < EP_VEC; q++) { const int i = tid + q * EP_THREADS; if (i < nvec) ((uint4 *)s_bytes)[i] = pv[q]; }
            if (c0 + EP_TILE < W_env) {
                int sh1, nv1;
                const uint4 *g1 = tile_src(c0 + EP_TILE, &sh1, &nv1);
#pragma unroll
                for (int q = 0; q < EP_VEC; q++) { const int i = tid + q * EP_THREADS; pv[q] = i < nv1 ? g1[i] : make_uint4(0, 0, 0, 0); }
            }
            __syncthreads();
        }
        // ---- this thread's EP_PER consecutive pixels ----
        const unsigned *wds = BULK ? (const unsigned *)(s_bytes + (size_t)(tile & 1) * EP_TILE * 4) + px0
                                   : (const unsigned *)s_bytes + (shift >> 2) + px0;
        double sx = 0, sy = 0, sY = 0, sw_ = 0;
        unsigned char *st = s_stage + (size_t)tid * (EP_PER * EP_ENTRY + 16);
#pragma unroll
        for (int k = 0; k < EP_PER; k++) {
            double *se = (double *)(st + k * EP_ENTRY);
            se[0] = sx; se[1] = sy; se[2] = sY;
            if (RR_PREF_N == 4) se[3] = sw_;
            if (px0 + k < n) {
                const unsigned pxw = wds[k];
                const double bb = rr_u8_unit((uint8_t)(pxw & 255u)), gg = rr_u8_unit((uint8_t)((pxw >> 8) & 255u)),
                             rr = rr_u8_unit((uint8_t)((pxw >> 16) & 255u));
                double x, y, Y;
                rr_env_xyY(bb, gg, rr, &x, &y, &Y);                 // my_utils.py:56-68, generator.py:408
                const double w = wv[k];
                sx += x * w; sy += y * w; sY += Y * w; sw_ += w;
            }
        }
        // ---- exclusive scan of the thread totals: shuffle scan inside the warp, then the warp totals ----
        double ix = sx, iy = sy, iY = sY, iw = sw_;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double ux = __shfl_up_sync(0xffffffffu, ix, o), uy = __shfl_up_sync(0xffffffffu, iy, o);
            double uY = __shfl_up_sync(0xffffffffu, iY, o), uw = __shfl_up_sync(0xffffffffu, iw, o);
            if (lane >= o) { ix += ux; iy += uy; iY += uY; iw += uw; }
        }
        if (lane == 31) { s_wtot[warp][0] = ix; s_wtot[warp][1] = iy; s_wtot[warp][2] = iY; s_wtot[warp][3] = iw; }
        double ex = __shfl_up_sync(0xffffffffu, ix, 1), ey = __shfl_up_sync(0xffffffffu, iy, 1);
        double eY = __shfl_up_sync(0xffffffffu, iY, 1), ew = __shfl_up_sync(0xffffffffu, iw, 1);
        if (lane == 0) { ex = ey = eY = ew = 0; }
        __syncthreads();
        double ox = cx, oy = cy, oY = cY, ow = cw, tx_ = cx, ty_ = cy, tY_ = cY, tw_ = cw;
#pragma unroll
        for (int k = 0; k < EP_WARPS; k++) {
            if (k == warp) { ox = tx_; oy = ty_; oY = tY_; ow = tw_; }
            tx_ += s_wtot[k][0]; ty_ += s_wtot[k][1]; tY_ += s_wtot[k][2]; tw_ += s_wtot[k][3];
        }
        s_toff[tid][0] = ox + ex; s_toff[tid][1] = oy + ey; s_toff[tid][2] = oY + eY; s_toff[tid][3] = ow + ew;
        cx = tx_; cy = ty_; cY = tY_; cw = tw_;
        __syncthreads();
        // ---- coalesced write-out: the tile's doubles in linear order (streaming stores: 1.1 GB per step that nothing re-reads soon) ----
        double *dst = prow + (size_t)c0 * RR_PREF_N;
        for (int h = tid; h < RR_PREF_N * n; h += EP_THREADS) {
            const int px = h / RR_PREF_N, comp = h - px * RR_PREF_N;
            const double v = *(const double *)(s_stage + (size_t)px * EP_ENTRY + (size_t)(px >> 3) * 16 + comp * 8);
            RR_STREAM_STORE(&dst[h], s_toff[px >> 3][comp] + v);
        }
    }
    if (tid == 0) {
        double *pe = prow + (size_t)W_env * RR_PREF_N;
        pe[0] = cx; pe[1] = cy; pe[2] = cY;
        if (RR_PREF_N == 4) pe[3] = cw;
        rowtot[(size_t)f * H + r] = cY;
    }
}

static cudaError_t launch_env_prefix(const uint8_t *env8, const double *omega, double *pref, double *rowtot, int F, int H, int W_env,
                                     int pitch, bool bulk, cudaStream_t st) {
    dim3 g3(H, F);
    if (bulk) k_env_prefix<true><<<g3, EP_THREADS, 0, st>>>(env8, omega, pref, rowtot, H, W_env, pitch);
    else k_env_prefix<false><<<g3, EP_THREADS, 0, st>>>(env8, omega, pref, rowtot, H, W_env, pitch);
    return cudaGetLastError();
}

// sum over the map of omega * Y per frame: one warp, lanes take rows in stride, then a fixed shuffle tree (deterministic)
__global__ void k_ambient(const double *rowtot, double *ambient, int H) {
    const int f = blockIdx.x;
    double s = 0;
    for (int r = threadIdx.x; r < H; r += 32) s += rowtot[(size_t)f * H + r];
    s = warp_sum(s);
    if (threadIdx.x == 0) ambient[f] = s;
}

cudaError_t rr_launch_env(const rr_frame_bufs &b, const rr_static_tabs &t, int F, int W, int H, int W_env, cudaStream_t st) {
    dim3 g2((W_env + ENV_TX - 1) / ENV_TX, (H + ENV_TY - 1) / ENV_TY, F);
    k_env_map<<<g2, 256, 0, st>>>(b.bg8, t.env_src, t.env_written, t.env_tile_hole, b.env8, H, W_env, b.env_pitch, W * H);
    cudaError_t e = launch_env_prefix(b.env8, t.omega, b.pref, b.rowtot, F, H, W_env, b.env_pitch, b.env_bulk != 0, st);
    if (e != cudaSuccess) return e;
    k_ambient<<<F, 32, 0, st>>>(b.rowtot, b.ambient, H);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// per streak: FOV polygon -> mask spans -> solid-angle weighted sums -> tint; patch plan
//   (bad_weather.py:596-704, 363-413; generator.py:119-171).  One warp per streak.
// ------------------------------------------------------------------------------------------
struct rr_plan;
__device__ __forceinline__ void plan_sizes(const rr_plan &p, long long *g, long long *v, long long *a, int *vx0, int *vw);
// Column intervals of one mask row, merged, handed to `emit(lo, hi)` left to right: what rr_fcp_row (rr_cvmath.h) returns,
// without its arrays.  A row of a (near-)convex polygon has the scan span and the Bresenham runs of the two or three outline
// edges that cross it: up to SETUP_NI intervals live in registers (slots filled by predicated moves, a fixed compare-exchange
// network sorts them, the merge walks the fixed slots); a row with more falls back to the general routine.  The general
// routine's insertion sort lives in local memory and ran with 13 of 32 lanes active (30 % of k_setup's stall samples).
#define SETUP_NI 6
template <class EMIT>
__device__ __forceinline__ void setup_row_intervals(const rr_fcp &f, int y, EMIT &&emit) {
    if (y < 0 || y >= f.H) return;
    int L[SETUP_NI], Hh[SETUP_NI];
#pragma unroll
    for (int k = 0; k < SETUP_NI; k++) { L[k] = 0x7fffffff; Hh[k] = -0x7fffffff; }
    int n = 0;
    auto push = [&](int a, int b) {
#pragma unroll
        for (int k = 0; k < SETUP_NI; k++) if (n == k) { L[k] = a; Hh[k] = b; }
        n++;
    };
    if (y >= f.ymin && y <= f.ymax && y < f.y_stop) {
        int64_t xa, xb;
        if (rr_fcp_side_x(f, 0, y, &xa) && rr_fcp_side_x(f, 1, y, &xb)) {
            if (xa > xb) { int64_t t = xa; xa = xb; xb = t; }
            const int64_t half = 1 << 15;
            int xx1 = (int)((xa + half) >> 16), xx2 = (int)((xb + half) >> 16);
            if (xx2 >= 0 && xx1 < f.W) {
                if (xx1 < 0) xx1 = 0;
                if (xx2 >= f.W) xx2 = f.W - 1;
                if (xx1 <= xx2) push(xx1, xx2);
            }
        }
    }
    for (int k = 0; k < f.ne; k++) {
        int a, b;
        if (rr_line_row(f.ex0[k], f.ey0[k], f.ex1[k], f.ey1[k], y, &a, &b)) push(a, b);
    }
    if (n == 0) return;
    if (n > SETUP_NI) {                       // rare: many outline runs on one row
        int lo[RR_MAX_POLY + 2], hi[RR_MAX_POLY + 2];
        const int m = rr_fcp_row(f, y, lo, hi);
        for (int j = 0; j < m; j++) emit(lo[j], hi[j]);
        return;
    }
    // sort the six slots by their left end (empty slots carry INT_MAX and sink to the end): odd-even transposition network
#define SETUP_CE(i, j) { if (L[i] > L[j]) { int t_ = L[i]; L[i] = L[j]; L[j] = t_; t_ = Hh[i]; Hh[i] = Hh[j]; Hh[j] = t_; } }
#pragma unroll
    for (int round = 0; round < SETUP_NI; round++) {
        if (round & 1) { SETUP_CE(1, 2) SETUP_CE(3, 4) }
        else { SETUP_CE(0, 1) SETUP_CE(2, 3) SETUP_CE(4, 5) }
    }
#undef SETUP_CE
    int cl = L[0], ch = Hh[0];
#pragma unroll
    for (int k = 1; k < SETUP_NI; k++) {
        if (L[k] != 0x7fffffff) {
            if (L[k] <= ch + 1) { if (Hh[k] > ch) ch = Hh[k]; }       // touching or overlapping: one interval
            else { emit(cl, ch); cl = L[k]; ch = Hh[k]; }
        }
    }
    emit(cl, ch);
}

#define SETUP_WARPS 4
#ifndef SETUP_MINB
#define SETUP_MINB 8          // 64 registers: latency bound, more resident warps win (sweep r01h)
#endif
// One warp per streak: the prepared polygon walker (k_plan) is copied into shared memory, the 32 lanes take the
// rows of the mask, merge the scan span with the Bresenham outline runs, and read pref[hi+1]-pref[lo].
__global__ void __launch_bounds__(SETUP_WARPS * 32, SETUP_MINB) k_setup(rr_frame_bufs b, rr_static_tabs t, rr_cam_dev cam, int F,
                                                             int n_streaks) {
    __shared__ __align__(16) rr_fcp s_fcp[SETUP_WARPS];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int s = blockIdx.x * SETUP_WARPS + warp;
    if (s >= n_streaks) return;
    // frame of this streak: binary search in offsets
    int lo = 0, hi = F;
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (b.offsets[mid] <= s) lo = mid; else hi = mid; }
    const int f = lo;
    rr_fcp &fc = s_fcp[warp];
    {
        const int4 *src = (const int4 *)(b.fcp + s);
        int4 *dst = (int4 *)&fc;
        for (int i = lane; i < (int)(sizeof(rr_fcp) / sizeof(int4)); i += 32) dst[i] = src[i];
    }
    __syncwarp();
    const int rows = cam.H_env, cols = cam.W_env;
    const int m = fc.npts;
    double sx = 0, sy = 0, sY = 0, sw = 0;
    if (m > 0) {
        int ymin = 0x7fffffff, ymax = -0x7fffffff;
        for (int i = 0; i < m; i++) { int vy = fc.vy[i]; ymin = vy < ymin ? vy : ymin; ymax = vy > ymax ? vy : ymax; }
        if (ymin < 0) ymin = 0;
        if (ymax > rows - 1) ymax = rows - 1;
        const size_t stride = (size_t)(cols + 1);
        const double *P = b.pref + (size_t)f * rows * stride * RR_PREF_N;
        for (int y = ymin + lane; y <= ymax; y += 32) {
            setup_row_intervals(fc, y, [&](int lo_, int hi_) {
                const double *a = P + ((size_t)y * stride + lo_) * RR_PREF_N, *e = P + ((size_t)y * stride + hi_ + 1) * RR_PREF_N;
                sx += e[0] - a[0]; sy += e[1] - a[1]; sY += e[2] - a[2];
                // the solid angles do not depend on the frame: their row prefix is a per-camera table (L2 resident)
                if (RR_PREF_N == 4) sw += e[3] - a[3];
                else sw += t.omega_pref[(size_t)y * stride + hi_ + 1] - t.omega_pref[(size_t)y * stride + lo_];
            });
        }
        sx = warp_sum(sx); sy = warp_sum(sy); sY = warp_sum(sY); sw = warp_sum(sw);
    }
    if (lane == 0) {
        // the geometric half of the plan was written by k_plan (which also withdrew degenerate streaks); add the photometry.
        // Only the tint / diagnostic fields are written: the streak chain may be reading the geometry fields right now.
        rr_plan &P = b.plans[s];
        const bool ok = P.valid && m > 0;
        if (ok) {
            double omega_total = *t.omega_total;
            double fov_x = sx / sw, fov_y = sy / sw;                        // bad_weather.py:397
            double ambient = b.ambient[f] / omega_total;                    // :403-404
            double avg_fov_lum = sY / omega_total;                          // :407
            double drop_Y = 0.94 * avg_fov_lum + 0.06 * ambient;            // :408
            P.fov_x = fov_x; P.fov_y = fov_y; P.drop_Y = drop_Y;
            double kb, kg, kr;
            rr_tint(fov_x, fov_y, drop_Y, &kb, &kg, &kr);
            P.kb = kb; P.kg = kg; P.kr = kr;
        }
    }
}

// Everything about a streak that needs nothing from the frame, one THREAD per streak (serial computations that would
// leave 31 lanes of k_setup's warp idle):
//   * the geometric half of the plan (patch warp, defocus, placement: rr_plan_patch -- an 8x8 LU for Big drops,
//     trigonometry for the others)
//   * the field-of-view polygon: 20 cone rays -> lat-long vertices, wrap splice, Clipper restatement, and the
//     prepared cv::fillConvexPoly walker (bad_weather.py:596-704, 363-373, 388), left in global memory for k_setup
__global__ void __launch_bounds__(128) k_plan(rr_frame_bufs b, rr_static_tabs t, rr_cam_dev cam, int n_streaks) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streaks) return;
    const rr_streak_rec rec = b.streaks[s];
    rr_plan p;
    memset(&p, 0, sizeof(p));
    bool ok = rec.tex_idx < cam.n_tex;
    if (ok) ok = rr_plan_patch(rec, cam, t.tex_h[rec.tex_idx], p);
    if (ok) { p.tex_off = t.tex_off[rec.tex_idx]; p.g_off = t.tex_poff[rec.tex_idx]; }
    else p.pw = p.ph = p.bw = p.bh = 0;
    // field-of-view mask
    __align__(16) rr_fcp fc;
    memset(&fc, 0, sizeof(fc));
    if (ok) {
        const int rows = cam.H_env, cols = cam.W_env;
        double px[RR_MAX_POLY], py[RR_MAX_POLY];
        const int npoly = rr_fov_polygon(rec, cam.radius, cam.fov_deg, rows, cols, px, py);
        const int m = npoly > 0 ? rr_clip_fov_polygon(px, py, npoly, cols, rows, fc.vx, fc.vy) : 0;
        fc.npts = m;
        if (m > 0) rr_fcp_prepare(fc, cols, rows);
        else ok = false;                                  // degenerate field of view: the reference skips the streak (generator.py:185-189)
    }
    // the streak's fate is settled here (k_setup only adds the photometry), so that the arena scan, the rasteriser and the
    // blur can run beside the frame stages
    if (!ok) p.pw = p.ph = p.bw = p.bh = 0;
    p.valid = ok ? 1 : 0;
    b.plans[s] = p;
    long long g_, v_, a_; int vx0_, vw_;
    plan_sizes(p, &g_, &v_, &a_, &vx0_, &vw_);
    b.sizes[s] = make_int4((int)g_, (int)v_, (int)a_, 0);
    b.boxes[s] = a_ > 0 ? make_int4(p.bx0, p.by0, p.bw, p.bh) : make_int4(0, 0, 0, 0);
    {
        const int4 *src = (const int4 *)&fc;
        int4 *dst = (int4 *)(b.fcp + s);
        for (int i = 0; i < (int)(sizeof(rr_fcp) / sizeof(int4)); i++) dst[i] = src[i];
    }
}

cudaError_t rr_launch_plan(const rr_frame_bufs &b, const rr_static_tabs &t, const rr_cam_dev &cam, int n_streaks, cudaStream_t st) {
    if (n_streaks == 0) return cudaSuccess;
    k_plan<<<(n_streaks + 127) / 128, 128, 0, st>>>(b, t, cam, n_streaks);
    return cudaGetLastError();
}

cudaError_t rr_launch_setup(const rr_frame_bufs &b, const rr_static_tabs &t, const rr_cam_dev &cam, int F, int n_streaks,
                            cudaStream_t st) {
    if (n_streaks == 0) return cudaSuccess;
    k_setup<<<(n_streaks + SETUP_WARPS - 1) / SETUP_WARPS, SETUP_WARPS * 32, 0, st>>>(b, t, cam, F, n_streaks);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// arena layout + work lists: exclusive scans over the streaks (single block, deterministic)
//   scan[i*6 + 0..2]: element offsets of g (pre-blur patch), v (column-pass result), a (blurred alpha block)
//   scan[n*6]: total elements (read back by the host when the arena overflows)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void blur_extents(const rr_plan &p, int *vx0, int *vw);
// arena elements of one streak: g (pre-blur patch), v (column pass), a (visible blurred block)
__device__ __forceinline__ void plan_sizes(const rr_plan &p, long long *g, long long *v, long long *a, int *vx0, int *vw) {
    *g = *v = *a = 0; *vx0 = 0; *vw = 0;
    if (p.valid && p.bw > 0 && p.bh > 0 && p.pw > 0 && p.ph > 0) {
        blur_extents(p, vx0, vw);
        *g = (long long)p.pw * p.ph; *v = (long long)(*vw) * p.bh; *a = (long long)p.bw * p.bh;
    }
}
__device__ __forceinline__ void blur_extents(const rr_plan &p, int *vx0, int *vw) {
    // padded columns of the column-pass result that the visible block needs and that are non-zero
    int a = p.cropx - p.rx, e = p.cropx + p.bw + p.rx;
    if (a < p.shift) a = p.shift;
    if (e > p.shift + p.pw) e = p.shift + p.pw;
    *vx0 = a;
    *vw = e > a ? e - a : 0;
}

__global__ void __launch_bounds__(1024) k_scan(rr_frame_bufs b, int n) {
    // one block: per-thread totals, a shuffle scan over the 1024 threads, then the per-streak offsets.
    // Integer sums: exact and order independent.
    __shared__ long long wtot[32];
    __shared__ long long grand;
    __shared__ int overflow;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (n + 1023) / 1024;
    const int i0 = tid * per, i1 = i0 + per < n ? i0 + per : n;
    long long t = 0;
    for (int i = i0; i < i1; i++) {
        int4 sz = b.sizes[i];
        t += (long long)sz.x + sz.y + sz.z;
    }
    long long inc = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { long long u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        long long v = wtot[lane], w = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { long long u = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += u; }
        wtot[lane] = w - v;                          // exclusive prefix of the warp totals
        if (lane == 31) grand = w;
    }
    __syncthreads();
    if (tid == 0) {
        overflow = grand > b.arena_cap;
        if (overflow) *b.err_flag = 1;                   // nothing is rendered: the host grows the arena and re-runs
        b.scan[(size_t)n * 6] = grand;
    }
    __syncthreads();
    long long el = wtot[warp] + inc - t;
    for (int i = i0; i < i1; i++) {
        int4 sz = b.sizes[i];
        long long g = sz.x, v = sz.y, a = sz.z;
        if (overflow) { b.plans[i].valid = 0; b.plans[i].bw = b.plans[i].bh = 0; b.boxes[i] = make_int4(0, 0, 0, 0); }
        long long *sc = b.scan + (size_t)i * 6;
        sc[0] = el; sc[1] = el + g; sc[2] = el + g + v;
        el += g + v + a;
    }
}

cudaError_t rr_launch_scan(const rr_frame_bufs &b, int n_streaks, cudaStream_t st) {
    k_scan<<<1, 1024, 0, st>>>(b, n_streaks);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// pre-blur gray patches  (generator.py:126-171): one CTA per streak (persistent, strided).
//   Big                -> one thread per patch pixel (16-tap bicubic perspective warp)
//   INTER_AREA shrink  -> cv::resizeArea_ order: for every canvas row and patch column dx the weighted
//                         sum of the row's canvas pixels (left to right), then the rows top to bottom.
//                         One thread owns one (canvas row, dx) chain and evaluates its canvas pixels
//                         (imutils.rotate_bound -> warpAffine fixed-point bilinear) on the fly, skipping
//                         the columns where the rotated texture cannot be (exact zeros); the chains of a
//                         whole canvas usually fit one band, so a streak costs three barriers.
//   integer-factor INTER_AREA (resizeAreaFast_): rare; its 4-wide groups run across rows, so the canvas
//                         is staged band by band in shared memory and every patch pixel walks its cell.
// ------------------------------------------------------------------------------------------
#define RAS_THREADS 128
#ifndef RAS_CAP
#define RAS_CAP 1280          // doubles per staging array (sweep r01h: 1024/6 CTAs 2.86 ms, 1280/5 2.05, 1792/4 2.16, 2304/3 2.47)
#endif
#ifndef RAS_MINB
#define RAS_MINB 5            // resident CTAs per SM the register allocation is tuned for
#endif
#ifndef RAS_PADDED
#define RAS_PADDED 1          // chain-loop sampler: 1 = zero-bordered texture copies (no per-tap predicates), 0 = plain textures
#endif
#if RAS_PADDED
#define RAS_SAMPLE(X, Y) ras_sample_pad(texp, tw, th, X, Y)
#else
#define RAS_SAMPLE(X, Y) ras_sample_int(tex, tw, th, X, Y)
#endif
#ifndef RAS_UNROLL
#define RAS_UNROLL 2          // canvas samples a chain keeps in flight
#endif
#define RAS_MAXW 512          // widest rotated canvas handled by the staged path
#define RAS_TXN 128           // widest / tallest patch with cached computeResizeAreaTab spans
#define RAS_MAXD 1024         // patch pixels with accumulators resident in shared memory
#define RAS_RBMAX 128         // rows per band of the area-fast path (RAS_CAP / canvas width; canvases are at least 32 wide)
#define RAS_SMEM_BYTES (sizeof(double) * (2 * RAS_CAP + RAS_MAXD + 256) + sizeof(rr_area_span) * 2 * RAS_TXN + \
                        sizeof(int) * (2 * RAS_MAXW + 2 * RAS_RBMAX))

// Bilinear weights of remapBilinear: w = (1 - fy/32 or fy/32) * (1 - fx/32 or fx/32) in float32.  With 5-bit fractions
// both factors and their product are exact, so w == p / 1024 with the integer p = {32 - fy, fy} * {32 - fx, fx}, and
// fl(v * w) == fl((v / 1024) * p) bit for bit (scaling by a power of two is exact): the kernel keeps the texture
// look-up table pre-scaled by 2^-10 (lut[u] = u / 255.0 / 1024) and multiplies by the integer products.
#define RAS_LUT_SCALE 0.0009765625
#define RAS_UNIT (1.0 / 261120.0)     // 1 / (255 * 1024): what one unit of ras_sample_int's integer sum is worth
__device__ __forceinline__ double ras_sample(const uint8_t *tex, int tw, int th, const double *lut, int X, int Y) {
    // rr_warp_affine_linear with the fixed-point coordinates already formed.  One code path for interior and
    // border samples (out-of-texture taps contribute the border value 0 with their weight, exactly the
    // expression remapBilinear evaluates), so a warp does not diverge along the texture outline.
    const int sx = X >> RR_INTER_BITS, sy = Y >> RR_INTER_BITS;     // |coordinates| << 2^15: the short saturation cannot act
    const bool x0ok = (unsigned)sx < (unsigned)tw, x1ok = (unsigned)(sx + 1) < (unsigned)tw;
    const bool y0ok = (unsigned)sy < (unsigned)th, y1ok = (unsigned)(sy + 1) < (unsigned)th;
    if (!((x0ok | x1ok) & (y0ok | y1ok))) return 0.0;               // all four taps outside: the border constant
    const int fx = X & (RR_INTER_TAB - 1), fy = Y & (RR_INTER_TAB - 1);
    const int gx = RR_INTER_TAB - fx, gy = RR_INTER_TAB - fy;
    const double w0 = (double)(gy * gx), w1 = (double)(gy * fx), w2 = (double)(fy * gx), w3 = (double)(fy * fx);
    const uint8_t *S = tex + sy * tw + sx;
    const double v0 = (x0ok & y0ok) ? lut[S[0]] : 0.0;
    const double v1 = (x1ok & y0ok) ? lut[S[1]] : 0.0;
    const double v2 = (x0ok & y1ok) ? lut[S[tw]] : 0.0;
    const double v3 = (x1ok & y1ok) ? lut[S[tw + 1]] : 0.0;
    return v0 * w0 + v1 * w1 + v2 * w2 + v3 * w3;
}

// The same sample as an exact integer: sum of texel * weight products, 0 .. 255 * 1024.  remapBilinear rounds each of its
// four float64 products; the integer sum carries no rounding at all, and the two differ by < 3e-16 relative -- nine
// orders of magnitude below the float32 ULP the output is held to.  It takes the dependent look-up-table loads, four
// int -> double conversions and seven float64 operations per canvas pixel out of the chain loop.
__device__ __forceinline__ int ras_sample_int(const uint8_t *tex, int tw, int th, int X, int Y) {
    const int sx = X >> RR_INTER_BITS, sy = Y >> RR_INTER_BITS;
    const bool x0ok = (unsigned)sx < (unsigned)tw, x1ok = (unsigned)(sx + 1) < (unsigned)tw;
    const bool y0ok = (unsigned)sy < (unsigned)th, y1ok = (unsigned)(sy + 1) < (unsigned)th;
    if (!((x0ok | x1ok) & (y0ok | y1ok))) return 0;
    const int fx = X & (RR_INTER_TAB - 1), fy = Y & (RR_INTER_TAB - 1);
    const uint8_t *S = tex + sy * tw + sx;
    const int t00 = (x0ok & y0ok) ? S[0] : 0, t01 = (x1ok & y0ok) ? S[1] : 0;
    const int t10 = (x0ok & y1ok) ? S[tw] : 0, t11 = (x1ok & y1ok) ? S[tw + 1] : 0;
    const int top = (RR_INTER_TAB - fx) * t00 + fx * t01, bot = (RR_INTER_TAB - fx) * t10 + fx * t11;
    return (RR_INTER_TAB - fy) * top + fy * bot;
}

// The chain loop's sampler on ZERO-BORDERED copies of the textures: a texture of h x w texels is stored as (h + 2) x (w + 2) bytes
// with a one-texel border of zeros (the warp's border constant), so the four taps of a bilinear sample at (sx, sy) with
// sx in [-1, w - 1], sy in [-1, h - 1] are four unconditional byte loads -- no per-tap predicates and selects; outside that
// range all four taps are border and the sample is 0.  Same bytes per texel as the plain copy (a 4-texel word per sample was
// measured slower: four times the footprint in L1).
__device__ __forceinline__ int ras_sample_pad(const uint8_t *P, int tw, int th, int X, int Y) {
    const int sx = X >> RR_INTER_BITS, sy = Y >> RR_INTER_BITS;
    if ((unsigned)(sx + 1) > (unsigned)tw || (unsigned)(sy + 1) > (unsigned)th) return 0;
    const uint8_t *S = P + (sy + 1) * (tw + 2) + (sx + 1);
    const int fx = X & (RR_INTER_TAB - 1), fy = Y & (RR_INTER_TAB - 1);
    const int top = (RR_INTER_TAB - fx) * (int)S[0] + fx * (int)S[1];
    const int bot = (RR_INTER_TAB - fx) * (int)S[tw + 2] + fx * (int)S[tw + 3];
    return (RR_INTER_TAB - fy) * top + fy * bot;
}

__global__ void k_build_padded(const uint8_t *db, const int32_t *tex_off, const int32_t *tex_h, const int32_t *tex_poff, int tw, uint8_t *dbp) {
    const int k = blockIdx.y, th = tex_h[k];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (th + 2) * (tw + 2)) return;
    const int r = i / (tw + 2), c = i - r * (tw + 2);
    const int y = r - 1, x = c - 1;
    dbp[tex_poff[k] + i] = (y >= 0 && y < th && x >= 0 && x < tw) ? db[tex_off[k] + y * tw + x] : (uint8_t)0;
}

cudaError_t rr_launch_build_padded(const uint8_t *db, const int32_t *tex_off, const int32_t *tex_h, const int32_t *tex_poff, int n_tex, int tw,
                                   int max_h, uint8_t *dbp, cudaStream_t st) {
    dim3 g(((max_h + 2) * (tw + 2) + 255) / 256, n_tex);
    k_build_padded<<<g, 256, 0, st>>>(db, tex_off, tex_h, tex_poff, tw, dbp);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(RAS_THREADS, RAS_MINB) k_raster(rr_frame_bufs b, rr_static_tabs t, rr_cam_dev cam, int n) {
    extern __shared__ double ras_smem[];
    double *C = ras_smem;                       // [RAS_CAP]   area-fast: canvas band;  area: first half of the chain sums
    double *BUF = C + RAS_CAP;                  // [RAS_CAP]   area-fast: quad accumulators; area: second half of the chain sums
    double *SUM = BUF + RAS_CAP;                // [RAS_MAXD]  per patch pixel running sum
    double *lut = SUM + RAS_MAXD;               // [256]       u8 / 255.0 / 1024 (see ras_sample)
    rr_area_span *TX = (rr_area_span *)(lut + 256);     // [RAS_TXN]
    rr_area_span *TY = TX + RAS_TXN;                     // [RAS_TXN]
    int *adx = (int *)(TY + RAS_TXN), *bdx = adx + RAS_MAXW;   // [RAS_MAXW] each
    int *XR = bdx + RAS_MAXW, *YR = XR + RAS_RBMAX;            // [RAS_RBMAX] each (area-fast)
    double *ACC = BUF;
    double *CB = C;                             // [2 * RAS_CAP] chain sums of the area path: CB[r * pw + dx]
    __shared__ rr_plan sp;
    const int tid = threadIdx.x;
    for (int i = tid; i < 256; i += RAS_THREADS) lut[i] = rr_u8_unit((uint8_t)i) * RAS_LUT_SCALE;     // bad_weather.py:252
    const int tw = cam.db_width;
    for (int s = blockIdx.x; s < n; s += gridDim.x) {
        __syncthreads();
        if (tid < sizeof(rr_plan) / 4) ((int *)&sp)[tid] = ((const int *)&b.plans[s])[tid];
        __syncthreads();
        const rr_plan &p = sp;
        long long g, vv, aa; int vx0_, vw_;
        plan_sizes(p, &g, &vv, &aa, &vx0_, &vw_);
        if (g == 0) continue;
        const uint8_t *tex = t.db + p.tex_off;
        const uint8_t *texp = t.dbp + p.g_off;          // the zero-bordered copy (ras_sample_pad)
        double *out = b.arena + b.scan[(size_t)s * 6 + 0];
        const bool staged = p.type != RR_BIG && (p.resize_mode == RR_RESIZE_AREA || p.resize_mode == RR_RESIZE_AREA_FAST) &&
                            p.nW <= RAS_MAXW && p.pw <= RAS_TXN && p.ph <= RAS_TXN && g <= RAS_MAXD;
        if (!staged) {
            for (int e = tid; e < (int)g; e += RAS_THREADS) {
                int y = e / p.pw, x = e - y * p.pw;
                out[e] = rr_patch_pixel(p, tex, tw, d_cubic, x, y);
            }
            continue;
        }
        const int nW = p.nW, nH = p.nH, pw = p.pw, ph = p.ph, th = p.tex_h, npx = pw * ph;
        const int AB_SCALE = 1 << 10;
        const bool fast = p.resize_mode == RR_RESIZE_AREA_FAST;
        for (int x = tid; x < nW; x += RAS_THREADS) {
            adx[x] = rr_round(p.M[0] * x * AB_SCALE);
            bdx[x] = rr_round(p.M[3] * x * AB_SCALE);
        }
        if (!fast) {
            for (int dx = tid; dx < pw; dx += RAS_THREADS) TX[dx] = rr_area_tab(dx, p.scale_x, nW);
            for (int dy = tid; dy < ph; dy += RAS_THREADS) TY[dy] = rr_area_tab(dy, p.scale_y, nH);
            const double M0 = p.M[0], M1 = p.M[1], M2 = p.M[2], M3 = p.M[3], M4 = p.M[4], M5 = p.M[5];
            const double inv0 = rr_canvas_inv(M0 * 1024.0), inv3 = rr_canvas_inv(M3 * 1024.0);
            int RB = (2 * RAS_CAP) / pw;             // pw <= RAS_TXN: at least 20 rows
            if (RB > nH) RB = nH;
            for (int s0 = 0; s0 < nH; s0 += RB) {
                const int rb = (nH - s0) < RB ? (nH - s0) : RB;
                __syncthreads();                     // tables ready / previous band's chain sums consumed
                // cv::resizeArea_: buf[dx] = sum_k S[sx_k] * alpha_k (left to right) for every source row of the band ...
                // consecutive threads take consecutive ROWS of the same patch column: their column ranges (and the
                // part of them the slanted texture covers) nearly coincide, so a warp's loops run in step
                for (int i = tid; i < rb * pw; i += RAS_THREADS) {
                    const int dx = i / rb, r = i - dx * rb;
                    const int sy = s0 + r;
                    const int yy = p.flip ? (nH - 1 - sy) : sy;
                    const int xr = rr_round((M1 * yy + M2) * AB_SCALE) + AB_SCALE / RR_INTER_TAB / 2;
                    const int yr = rr_round((M4 * yy + M5) * AB_SCALE) + AB_SCALE / RR_INTER_TAB / 2;
                    const rr_area_span tx = TX[dx];
                    // canvas columns of this chain: [first partial] s_first .. s_first + n - 1 [last partial] ...
                    int c_lo = tx.s_first - tx.has_first, c_hi = tx.s_first + tx.n - 1 + tx.has_last;
                    // ... of which only the slanted band the rotated texture covers can be non-zero; a zero sample adds
                    // exactly +0.0 to the running sum, so skipping it leaves the sum bit for bit unchanged
                    int z_lo, z_hi;
                    rr_canvas_row_span(p.M, inv0, inv3, xr, yr, nW, tw, th, &z_lo, &z_hi);
                    if (c_lo < z_lo) c_lo = z_lo;
                    if (c_hi > z_hi) c_hi = z_hi;
                    double buf = 0;
                    if (c_lo <= c_hi) {
                        // integer sample sums per weight class: the partial first / last column and the full-weight middle
                        // (at most 512 columns of at most 255 * 1024 each: no overflow)
                        const int m_lo = tx.s_first, m_hi = tx.s_first + tx.n;       // full-weight columns [m_lo, m_hi)
                        int sF = 0, sM = 0, sL = 0;
                        if (c_lo < m_lo) sF = RAS_SAMPLE((xr + adx[c_lo]) >> (10 - RR_INTER_BITS), (yr + bdx[c_lo]) >> (10 - RR_INTER_BITS));
                        if (c_hi >= m_hi) sL = RAS_SAMPLE((xr + adx[c_hi]) >> (10 - RR_INTER_BITS), (yr + bdx[c_hi]) >> (10 - RR_INTER_BITS));
                        const int a = c_lo < m_lo ? m_lo : c_lo, e = c_hi >= m_hi ? m_hi - 1 : c_hi;
                        int c = a;
                        for (; c + RAS_UNROLL - 1 <= e; c += RAS_UNROLL) {       // RAS_UNROLL independent samples in flight
                            int v[RAS_UNROLL];
#pragma unroll
                            for (int u = 0; u < RAS_UNROLL; u++)
                                v[u] = RAS_SAMPLE((xr + adx[c + u]) >> (10 - RR_INTER_BITS), (yr + bdx[c + u]) >> (10 - RR_INTER_BITS));
#pragma unroll
                            for (int u = 0; u < RAS_UNROLL; u++) sM += v[u];
                        }
                        for (; c <= e; c++) sM += RAS_SAMPLE((xr + adx[c]) >> (10 - RR_INTER_BITS), (yr + bdx[c]) >> (10 - RR_INTER_BITS));
                        // texel / 255 / 1024 folded into one constant (RAS_UNIT), the three weights applied left to right
                        buf = ((double)sF * RAS_UNIT) * (double)tx.a_first;
                        buf += ((double)sM * RAS_UNIT) * (double)tx.a_mid;
                        buf += ((double)sL * RAS_UNIT) * (double)tx.a_last;
                    }
                    CB[r * pw + dx] = buf;
                }
                __syncthreads();
                // ... then sum[dx] (+)= beta * buf[dx], rows top to bottom; a patch pixel's rows may span bands
                for (int e = tid; e < npx; e += RAS_THREADS) {
                    int dy = e / pw, dx = e - dy * pw;
                    const rr_area_span ty = TY[dy];
                    const int nr = ty.has_first + ty.n + ty.has_last, row0 = ty.s_first - ty.has_first;
                    int jlo = s0 - row0; if (jlo < 0) jlo = 0;
                    int jhi = s0 + rb - row0; if (jhi > nr) jhi = nr;
                    if (jlo >= jhi) continue;
                    double acc = SUM[e];
                    for (int j = jlo; j < jhi; j++) {
                        float beta = (ty.has_first && j == 0) ? ty.a_first : ((ty.has_last && j == nr - 1) ? ty.a_last : ty.a_mid);
                        double v = beta * CB[(row0 + j - s0) * pw + dx];
                        acc = (j == 0) ? v : acc + v;
                    }
                    SUM[e] = acc;
                }
            }
            __syncthreads();
            for (int e = tid; e < npx; e += RAS_THREADS) {
                double v = SUM[e];
                out[e] = v < 0 ? 0 : (v > 1 ? 1 : v);
            }
            continue;
        }
        // ---- integer-factor area mode: the canvas staged band by band; every canvas pixel is sampled exactly once ----
        const int isx = rr_round(p.scale_x), isy = rr_round(p.scale_y);
        const int area = isx * isy, area4 = area - (area & 3);
        int RB = RAS_CAP / (nW > pw ? nW : pw);
        if (RB < 1) RB = 1;
        if (RB > RAS_RBMAX) RB = RAS_RBMAX;
        const int step_r = RAS_THREADS / nW, step_c = RAS_THREADS - step_r * nW;
        const int r_first = tid / nW, c_first = tid - r_first * nW;
        for (int s0 = 0; s0 < nH; s0 += RB) {
            const int rb = (nH - s0) < RB ? (nH - s0) : RB;
            __syncthreads();                         // previous band fully consumed (C, ACC, XR)
            for (int r = tid; r < rb; r += RAS_THREADS) {
                int sy = s0 + r;
                int yy = p.flip ? (nH - 1 - sy) : sy;
                XR[r] = rr_round((p.M[1] * yy + p.M[2]) * AB_SCALE) + AB_SCALE / RR_INTER_TAB / 2;
                YR[r] = rr_round((p.M[4] * yy + p.M[5]) * AB_SCALE) + AB_SCALE / RR_INTER_TAB / 2;
            }
            __syncthreads();
            {
                int r = r_first, c = c_first;                       // (row, column) of flattened index tid
                for (int i = tid; i < rb * nW; i += RAS_THREADS) {
                    int X = (XR[r] + adx[c]) >> (10 - RR_INTER_BITS);
                    int Y = (YR[r] + bdx[c]) >> (10 - RR_INTER_BITS);
                    C[i] = ras_sample(tex, tw, th, lut, X, Y);
                    c += step_c; r += step_r;                       // advance by RAS_THREADS without a division
                    if (c >= nW) { c -= nW; r++; }
                }
            }
            __syncthreads();
            // cv::resizeAreaFast_: sum += ((a + b) + c) + d over the cell in row-major order, then the tail
            for (int e = tid; e < npx; e += RAS_THREADS) {
                int dy = e / pw, dx = e - dy * pw;
                const int row0 = dy * isy;
                int jlo = s0 - row0; if (jlo < 0) jlo = 0;
                int jhi = s0 + rb - row0; if (jhi > isy) jhi = isy;
                if (jlo >= jhi) continue;
                double sum = (jlo == 0) ? 0.0 : SUM[e], acc = (jlo == 0) ? 0.0 : ACC[RAS_CAP - 1 - e];
                for (int j = jlo; j < jhi; j++) {
                    const int kbase = j * isx;
                    const double *row = C + (row0 + j - s0) * nW + dx * isx;
                    for (int c = 0; c < isx; c++) {
                        int k = kbase + c;
                        double v = row[c];
                        if (k < area4) {
                            int pos = k & 3;
                            acc = pos == 0 ? v : acc + v;
                            if (pos == 3) sum += acc;
                        } else sum += v;
                    }
                }
                SUM[e] = sum; ACC[RAS_CAP - 1 - e] = acc;
            }
        }
        __syncthreads();
        for (int e = tid; e < npx; e += RAS_THREADS) {
            float scale = 1.f / area;
            double v = SUM[e] * scale;
            out[e] = v < 0 ? 0 : (v > 1 ? 1 : v);
        }
    }
}

cudaError_t rr_launch_raster(const rr_frame_bufs &b, const rr_static_tabs &t, const rr_cam_dev &cam, int n_streaks, int n_sm,
                             cudaStream_t st) {
    if (n_streaks == 0) return cudaSuccess;
    const size_t smem = RAS_SMEM_BYTES;
    int grid = n_sm * RAS_MINB;
    if (grid > n_streaks) grid = n_streaks;
    k_raster<<<grid, RAS_THREADS, smem, st>>>(b, t, cam, n_streaks);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// defocus: scipy.ndimage.gaussian_filter(drop, [c, c/2, 0]) on the zero-padded patch
// (bad_weather.py:286-298).  Only the alpha channel is filtered: before the blur the three
// colour channels equal alpha * k_c (tint of a gray texture), so blur(colour_c) = k_c * blur(alpha)
// up to float64 rounding (DESIGN.md "linear tint").
// ------------------------------------------------------------------------------------------
// One CTA per streak (persistent, strided): both passes of the streak back to back, its two weight tables built
// once.  The column-pass result goes through the arena (written and read by the same CTA, a barrier in between).
#define BLUR_THREADS 128
__global__ void __launch_bounds__(BLUR_THREADS) k_blur(rr_frame_bufs b, int n) {
    __shared__ double wy[2 * RR_MAX_GAUSS_R + 1], wx[2 * RR_MAX_GAUSS_R + 1];
    __shared__ double s_norm[2];
    const int tid = threadIdx.x;
    __shared__ rr_plan sp;
    for (int s = blockIdx.x; s < n; s += gridDim.x) {
        __syncthreads();                                         // previous streak done with the plan and the tables
        if (tid < (int)(sizeof(rr_plan) / 4)) ((int *)&sp)[tid] = ((const int *)&b.plans[s])[tid];      // one coalesced read
        __syncthreads();
        const rr_plan &p = sp;
        long long gg, nv, na; int vx0, vw;
        plan_sizes(p, &gg, &nv, &na, &vx0, &vw);
        if (na == 0) continue;                                   // block-uniform
        const int pw = p.pw, ph = p.ph, ry = p.ry, rx = p.rx, cropy = p.cropy, cropx = p.cropx, shift = p.shift, bw = p.bw;
        // rr_gauss_weights (SciPy _gaussian_kernel1d) for both axes, the exponentials spread over the block; the
        // normalising sums keep numpy's order (one thread each, in different warps)
        {
            const int ny = 2 * ry + 1, nx = 2 * rx + 1;
            const double fy = -0.5 / (p.sig_y * p.sig_y), fx = -0.5 / (p.sig_x * p.sig_x);
            for (int i = tid; i < ny; i += BLUR_THREADS) wy[i] = ry > 0 ? exp(fy * (double)((i - ry) * (i - ry))) : 1.0;
            for (int i = tid; i < nx; i += BLUR_THREADS) wx[i] = rx > 0 ? exp(fx * (double)((i - rx) * (i - rx))) : 1.0;
            __syncthreads();
            if (tid == 0) s_norm[0] = rr_np_sum_small(wy, ny);
            if (tid == 32) s_norm[1] = rr_np_sum_small(wx, nx);
            __syncthreads();
            const double sy = s_norm[0], sx = s_norm[1];
            for (int i = tid; i < ny; i += BLUR_THREADS) wy[i] = wy[i] / sy;
            for (int i = tid; i < nx; i += BLUR_THREADS) wx[i] = wx[i] / sx;
            __syncthreads();
        }
        const double *g = b.arena + b.scan[(size_t)s * 6 + 0];
        double *v = b.arena + b.scan[(size_t)s * 6 + 1];
        double *aout = b.arena + b.scan[(size_t)s * 6 + 2];
        // column pass (axis 0, sigma c): SciPy correlate1d, symmetric weights: centre first, then pairs from the outside in
        for (int e = tid; e < (int)nv; e += BLUR_THREADS) {
            int yy = e / vw, xx = e - yy * vw;
            int gx = vx0 + xx - shift;       // patch column
            int gy = cropy + yy - shift;     // patch row of the centre tap (may be outside: zero padding)
            double c = (gy >= 0 && gy < ph) ? g[gy * pw + gx] : 0.0;
            double tmp = c * wy[ry];
            for (int jj = -ry; jj < 0; jj++) {
                int ya = gy + jj, yb = gy - jj;
                double va = (ya >= 0 && ya < ph) ? g[ya * pw + gx] : 0.0;
                double vb = (yb >= 0 && yb < ph) ? g[yb * pw + gx] : 0.0;
                tmp += (va + vb) * wy[jj + ry];
            }
            v[e] = tmp;
        }
        __syncthreads();                                         // the CTA's own global writes are visible to it after the barrier
        // row pass (axis 1, sigma c / 2) over the visible block only
        for (int e = tid; e < (int)na; e += BLUR_THREADS) {
            int yy = e / bw, xx = e - yy * bw;
            const double *vr = v + yy * vw;
            int xc = cropx + xx - vx0;        // column of the centre tap in the column-pass result
            double c = (xc >= 0 && xc < vw) ? vr[xc] : 0.0;
            double tmp = c * wx[rx];
            for (int jj = -rx; jj < 0; jj++) {
                int xa = xc + jj, xb = xc - jj;
                double va = (xa >= 0 && xa < vw) ? vr[xa] : 0.0;
                double vb = (xb >= 0 && xb < vw) ? vr[xb] : 0.0;
                tmp += (va + vb) * wx[jj + rx];
            }
            aout[e] = tmp;
        }
    }
}

cudaError_t rr_launch_blur(const rr_frame_bufs &b, int n_streaks, int n_sm, cudaStream_t st) {
    if (n_streaks == 0) return cudaSuccess;
    int grid = n_sm * 16;
    if (grid > n_streaks) grid = n_streaks;
    k_blur<<<grid, BLUR_THREADS, 0, st>>>(b, n_streaks);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// ordered compositing  (bad_weather.py:429-460): the blend is streak-order dependent, so each
// pixel walks the streaks that cover its tile in record (XML) order.  No atomics.
// ------------------------------------------------------------------------------------------
// A CTA owns a region of 32 x (RR_COMP_WARPS * RR_COMP_PY) pixels, each warp a strip of 32 x RR_COMP_PY of it (a lane
// = one column).  Two stages per round of COMP_ROUND streaks: (1) every thread tests two streak boxes against the
// REGION and the hits are compacted, in record order, into a shared list (one barrier pair per round; a frame of
// the headline workload is one round); (2) each warp walks the list 32 entries at a time, tests them against its
// own strip, and applies the hits in ballot order = record order.
#define COMP_THREADS (RR_COMP_WARPS * 32)
#define COMP_ROUND (2 * COMP_THREADS)
#ifndef RR_COMP_MINB
#define RR_COMP_MINB 4         // sweep r01h: (rows per lane, CTAs per SM) (2,4) 0.558 ms, (4,4) 0.560, (4,3) 0.583, (4,2) 0.673, (8,2) 0.767
#endif
__global__ void __launch_bounds__(COMP_THREADS, RR_COMP_MINB) k_composite(rr_frame_bufs b, rr_cam_dev cam, int tiles_x, int n_partials) {
    __shared__ int4 s_box[COMP_ROUND];
    __shared__ int s_idx[COMP_ROUND];
    __shared__ int s_wcnt[2][RR_COMP_WARPS];
    const int f = blockIdx.z;
    const int W = cam.W, H = cam.H;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * 32, x = tx0 + lane;
    const int ry0 = blockIdx.y * (RR_COMP_WARPS * RR_COMP_PY);           // region rows [ry0, ry0 + RR_COMP_WARPS * RR_COMP_PY)
    const int strip = blockIdx.y * RR_COMP_WARPS + warp, y0 = strip * RR_COMP_PY;
    const size_t npix = (size_t)W * H;
    double vb[RR_COMP_PY], vg[RR_COMP_PY], vr[RR_COMP_PY], mask[RR_COMP_PY];
    bool touched[RR_COMP_PY];
    double *rb = b.rainy + ((size_t)f * 3 + 0) * npix, *rg = rb + npix, *rr = rg + npix;
#pragma unroll
    for (int k = 0; k < RR_COMP_PY; k++) {
        const bool inside = x < W && y0 + k < H;
        const size_t pix = (size_t)(y0 + k) * W + x;
        vb[k] = inside ? rb[pix] : 0.0; vg[k] = inside ? rg[pix] : 0.0; vr[k] = inside ? rr[pix] : 0.0;
        mask[k] = 0; touched[k] = false;
    }
    const int s0 = b.offsets[f], s1 = b.offsets[f + 1];
    const double exposure = cam.exposure_blend;
    for (int base = s0; base < s1; base += COMP_ROUND) {
        // ---- stage 1: region hits of this round, compacted in record order ----
        int4 box[2];
        unsigned bal[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int s = base + h * COMP_THREADS + tid;
            box[h] = make_int4(0, 0, 0, 0);
            if (s < s1) box[h] = b.boxes[s];                // (bx0, by0, bw, bh), bw = 0 for streaks that draw nothing
            const bool hit = box[h].z > 0 && box[h].w > 0 && box[h].x < tx0 + 32 && box[h].x + box[h].z > tx0 &&
                             box[h].y < ry0 + RR_COMP_WARPS * RR_COMP_PY && box[h].y + box[h].w > ry0;
            bal[h] = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) s_wcnt[h][warp] = __popc(bal[h]);
        }
        __syncthreads();
        int total = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            int pre = total;
#pragma unroll
            for (int k = 0; k < RR_COMP_WARPS; k++) {
                const int c = s_wcnt[h][k];
                if (k < warp) pre += c;
                total += c;
            }
            if (bal[h] & (1u << lane)) {
                const int slot = pre + __popc(bal[h] & ((1u << lane) - 1));
                s_box[slot] = box[h];
                s_idx[slot] = base + h * COMP_THREADS + tid;
            }
        }
        __syncthreads();
        // ---- stage 2: this warp's strip ----
        if (y0 < H) {
            for (int j0 = 0; j0 < total; j0 += 32) {
                const int j = j0 + lane;
                int4 bx = make_int4(0, 0, 0, 0);
                if (j < total) bx = s_box[j];
                const bool hit = j < total && bx.y < y0 + RR_COMP_PY && bx.y + bx.w > y0;
                unsigned hb = __ballot_sync(0xffffffffu, hit);
                while (hb) {
                    const int k = __ffs(hb) - 1;
                    hb &= hb - 1;
                    const int sk = s_idx[j0 + k];
                    const int bx0 = __shfl_sync(0xffffffffu, bx.x, k), by0 = __shfl_sync(0xffffffffu, bx.y, k);
                    const int bw = __shfl_sync(0xffffffffu, bx.z, k), bh = __shfl_sync(0xffffffffu, bx.w, k);
                    const rr_plan *pp = b.plans + sk;        // same address in every lane: one broadcast transaction
                    const double kb = pp->kb, kg = pp->kg, kr = pp->kr, tau_one = pp->a_scale, c_scale = pp->c_scale;
                    const double *A = b.arena + b.scan[(size_t)sk * 6 + 2];
                    const int lx = x - bx0;
#pragma unroll
                    for (int q = 0; q < RR_COMP_PY; q++) {
                        const int ly = y0 + q - by0;
                        if (lx >= 0 && lx < bw && ly >= 0 && ly < bh && x < W) {
                            const double a = A[(size_t)ly * bw + lx];
                            if (a == 0.0) continue;          // keep = 1, colour + 0, mask + 0: bit for bit a no-op
                            const double keep = 1. - ((a * tau_one) / exposure);         // bad_weather.py:443
                            const double nb = (keep * vb[q]) + (kb * a) * c_scale;
                            const double ng = (keep * vg[q]) + (kg * a) * c_scale;
                            const double nr = (keep * vr[q]) + (kr * a) * c_scale;
                            vb[q] = nb < 0 ? 0 : (nb > 1 ? 1 : nb);                      // :446
                            vg[q] = ng < 0 ? 0 : (ng > 1 ? 1 : ng);
                            vr[q] = nr < 0 ? 0 : (nr > 1 ? 1 : nr);
                            mask[q] += a;                                                // :450
                            touched[q] = true;
                        }
                    }
                }
            }
        }
        if (base + COMP_ROUND < s1) __syncthreads();         // the lists are rewritten by the next round
    }
    double part = 0, mlo = 1e300, mhi = -1e300;
#pragma unroll
    for (int k = 0; k < RR_COMP_PY; k++) {
        const bool inside = x < W && y0 + k < H;
        if (inside) {
            const size_t pix = (size_t)(y0 + k) * W + x;
            if (touched[k]) { rb[pix] = vb[k]; rg[pix] = vg[k]; rr[pix] = vr[k]; }       // untouched pixels keep the fogged value
            b.maskd[(size_t)f * npix + pix] = mask[k];
            mlo = mask[k] < mlo ? mask[k] : mlo; mhi = mask[k] > mhi ? mask[k] : mhi;
            part += (vb[k] + vg[k]) + vr[k];
        }
    }
    part = warp_sum(part);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double a = __shfl_xor_sync(0xffffffffu, mlo, o), c = __shfl_xor_sync(0xffffffffu, mhi, o);
        mlo = a < mlo ? a : mlo; mhi = c > mhi ? c : mhi;
    }
    if (lane == 0) {
        const size_t slot = (size_t)f * n_partials + (size_t)strip * tiles_x + blockIdx.x;
        b.tile_sum[slot] = part; b.tile_min[slot] = mlo; b.tile_max[slot] = mhi;
    }
}

__global__ void k_frame_mean(rr_frame_bufs b, int n_tiles, int stride, double npix3) {
    int f = blockIdx.x;
    __shared__ double sh[256], shlo[256], shhi[256];
    double s = 0, lo = 1e300, hi = -1e300;
    for (int i = threadIdx.x; i < n_tiles; i += 256) {
        s += b.tile_sum[(size_t)f * stride + i];
        const double a = b.tile_min[(size_t)f * stride + i], c = b.tile_max[(size_t)f * stride + i];
        lo = a < lo ? a : lo; hi = c > hi ? c : hi;                 // strips below the image hold the neutral elements
    }
    sh[threadIdx.x] = s; shlo[threadIdx.x] = lo; shhi[threadIdx.x] = hi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sh[threadIdx.x] += sh[threadIdx.x + o];
            shlo[threadIdx.x] = shlo[threadIdx.x + o] < shlo[threadIdx.x] ? shlo[threadIdx.x + o] : shlo[threadIdx.x];
            shhi[threadIdx.x] = shhi[threadIdx.x + o] > shhi[threadIdx.x] ? shhi[threadIdx.x + o] : shhi[threadIdx.x];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        b.mask_range[2 * f] = shlo[0]; b.mask_range[2 * f + 1] = shhi[0];          // np.min / np.max of rainy_mask (exact, order free)
        double mean_rainy = sh[0] / npix3;                                          // generator.py:461
        double mean_bg = b.bgf ? ((b.bg_sum[f * 4] + b.bg_sum[f * 4 + 1]) + b.bg_sum[f * 4 + 2]) / npix3
                               : ((double)(b.chan_sum[f * 4] + b.chan_sum[f * 4 + 1] + b.chan_sum[f * 4 + 2]) / 255.0) / npix3;   // :462
        b.frame_mean[f] = mean_rainy - mean_bg;                                     // :463
    }
}

cudaError_t rr_launch_composite(const rr_frame_bufs &b, const rr_cam_dev &cam, int F, cudaStream_t st) {
    const int tiles_x = (cam.W + 31) / 32, ctas_y = (cam.H + RR_COMP_PY * RR_COMP_WARPS - 1) / (RR_COMP_PY * RR_COMP_WARPS);
    const int n_partials = (int)rr_n_partials(cam.W, cam.H);
    dim3 grid(tiles_x, ctas_y, F);
    k_composite<<<grid, RR_COMP_WARPS * 32, 0, st>>>(b, cam, tiles_x, n_partials);
    k_frame_mean<<<F, 256, 0, st>>>(b, tiles_x * ctas_y * RR_COMP_WARPS, n_partials, 3.0 * cam.W * cam.H);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// epilogue: mean shift, float32 / uint8 outputs  (generator.py:460-466)
// ------------------------------------------------------------------------------------------
__global__ void k_epilogue(rr_frame_bufs b, int npix, int F) {
    const int f = blockIdx.y;
    const size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= (size_t)npix) return;
    const size_t i = (size_t)f * npix + pix;
    double d = b.frame_mean[f];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        double v = b.rainy[((size_t)f * 3 + c) * npix + pix] - d;                   // :464
        if (b.out_bgr) RR_STREAM_STORE(&b.out_bgr[i * 3 + c], (float)v);
        if (b.out_u8) {
            double cl = v < 0 ? 0 : (v > 1 ? 1 : v);
            b.out_u8[i * 3 + c] = (uint8_t)(cl * 255);                              // plt.imsave float -> uint8
        }
    }
    // the rain mask in the forms a caller saves: float32, or what plt.imsave(path, rainy_mask) (generator.py:467) makes of
    // the float64 array -- Normalize: (m - min) / (max - min), zeros when flat; Colormap.__call__: int(t * 256), 256 -> 255
    const double m = b.maskd[i];
    if (b.out_mask) RR_STREAM_STORE(&b.out_mask[i], (float)m);
    if (b.out_idx8 || b.out_u16) {
        const double lo = b.mask_range[2 * f], hi = b.mask_range[2 * f + 1];
        double t = 0.0;
        if (hi > lo && m != lo) t = (m - lo) / (hi - lo);                           // m == lo: exactly 0 (and no zero-numerator division)
        if (b.out_idx8) { const double q = t * 256.0; b.out_idx8[i] = q >= 256.0 ? (uint8_t)255 : (uint8_t)(int)q; }
        if (b.out_u16) b.out_u16[i] = (uint16_t)(unsigned)(t * 65535.0 + 0.5);
    }
}

cudaError_t rr_launch_epilogue(const rr_frame_bufs &b, int F, int W, int H, cudaStream_t st) {
    dim3 grid((unsigned)(((size_t)W * H + 255) / 256), F);
    k_epilogue<<<grid, 256, 0, st>>>(b, W * H, F);
    return cudaGetLastError();
}

cudaError_t rr_launch_env_prefix_only(const rr_frame_bufs &b, const rr_static_tabs &t, int F, int H, int W_env, cudaStream_t st) {
    cudaError_t e = launch_env_prefix(b.env8, t.omega, b.pref, b.rowtot, F, H, W_env, b.env_pitch, b.env_bulk != 0, st);
    if (e != cudaSuccess) return e;
    k_ambient<<<F, 32, 0, st>>>(b.rowtot, b.ambient, H);
    return cudaGetLastError();
}


// The opt-in to more than 48 KB of dynamic shared memory is a per-DEVICE function attribute: every context sets it for
// its own device in rr_create (a process-wide "done once" flag would leave the second GPU of a process without it).
cudaError_t rr_prepare_device() {
    cudaError_t e = cudaFuncSetAttribute(k_fog<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(FOG_BYTES_A + FOG_BYTES_B + FOG_TY * FOG_TX * 3));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_fog<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(FOG_BYTES_A_TMA + FOG_BYTES_B + FOG_TY * FOG_TX * 3));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_fog_roll, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FOGR_SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_raster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RAS_SMEM_BYTES);
}

