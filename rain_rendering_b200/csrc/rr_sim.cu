// On-the-fly particle simulation: B200-native stand-in for the closed AHLSimulation binary
// (SURVEY.md 2.3).  What is restated from the binary (symbols / rodata):
//   * drop-size distribution: Marshall-Palmer N(D) = 8000 exp(-L D[mm]), L = 4.1 R^-0.21, inverse-CDF LUT
//   * emission: diameters drawn from N(D), water budget R[mm/h] * A / 3600 litres per second
//   * integrator: semi-implicit Euler at sim_hz, vertical forces only, a = -(m g - F_drag(v)) / m,
//     "constant speed" once a >= -0.1; dt kept in float32
//   * drag: F = 3 pi 1.8e-5 D v (1 + 0.16 Re^(2/3)) (1 + 0.013 (2.28 + We)^2.12 - 0.0746045)
//   * imaging: snapshots at shutter open / close, pinhole projection, kept iff an end point is inside
//     the sensor and the streak is not sub-pixel ("fog-like").
// What is NOT the binary's procedure: instead of time-stepping every drop of a large emitter box from a
// 5 s warm-up, the stationary drop field the warm-up converges to is sampled directly per frame
// (Poisson count, diameters weighted by 1/v and by the volume in which a drop of that size can be
// wider than min_width_px), and only the exposure interval is integrated with the binary's stepper.
// Frames are therefore independent, like everything downstream (common/generator.py:318).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <vector>
#include "rr_types.h"

#define SIM_NLUT 4096
#define SIM_PI 3.14159265358979323846

struct sim_consts {
    double f_px, W, H, T, dt;
    int n_steps;
    double z0, z1, wmin, vcam;
    double dmin, dmax;       // metres
    double margin_x;         // px
    uint64_t seed;
};

__host__ __device__ inline double sim_drag(double D, double v) {
    const double rho_air = 1.2047, mu = 1.8e-5, sigma = 0.073;
    double Re = (rho_air / mu) * D * v;
    double We = (rho_air / sigma) * D * v * v;
    return 3 * SIM_PI * mu * D * v * (1 + 0.16 * pow(Re, 2.0 / 3.0)) * (1 + 0.013 * pow(2.28 + We, 2.12) - 0.0746045);
}
__host__ __device__ inline double sim_mass(double D) { return 1000.0 * (4.0 / 3.0) * SIM_PI * (D / 2) * (D / 2) * (D / 2); }

// terminal velocity of the binary's own force model: m g = F_drag(v)
__host__ __device__ inline double sim_v_terminal(double D) {
    double mg = sim_mass(D) * 9.81;
    double lo = 0, hi = 40;
    for (int i = 0; i < 80; i++) {
        double mid = 0.5 * (lo + hi);
        if (sim_drag(D, mid) < mg) lo = mid; else hi = mid;
    }
    return 0.5 * (lo + hi);
}

__host__ __device__ inline uint64_t sim_mix(uint64_t x) {   // splitmix64 finaliser (counter-based RNG)
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ inline double sim_u01(uint64_t seed, uint64_t frame, uint64_t idx, uint64_t stream) {
    uint64_t h = sim_mix(seed ^ sim_mix(frame * 0x100000001B3ull + stream) ^ sim_mix(idx * 0xD6E8FEB86659FD93ull + 0x51ED270B7ull * stream));
    return ((double)(h >> 11) + 0.5) / 9007199254740992.0;
}

// volume (m^3) in which a drop of diameter D can be imaged wider than wmin, including the band above the
// sensor from which it falls into view during the exposure
__host__ __device__ inline double sim_region(const sim_consts &c, double D, double v, double *zmax_out) {
    double zmax = D * c.f_px / c.wmin;
    if (zmax > c.z1) zmax = c.z1;
    *zmax_out = zmax;
    if (zmax <= c.z0) return 0.0;
    // cross-section at depth z: (W + 2 mx) z / f  by  (H z / f + v T)
    double a = (c.W + 2 * c.margin_x) * c.H / (c.f_px * c.f_px);      // z^2 term
    double b = (c.W + 2 * c.margin_x) / c.f_px * (v * c.T);          // z term
    return a * (zmax * zmax * zmax - c.z0 * c.z0 * c.z0) / 3 + b * (zmax * zmax - c.z0 * c.z0) / 2;
}

// One candidate drop of one frame -> its imaged streak, or false when it is not imaged (outside the sensor, sub-pixel)
// (lut_v: terminal velocity of every table diameter, computed once per parameter set on the device -- k_sim_vt)
__device__ bool sim_candidate(const sim_consts &c, const double *cdf, const double *lut_d, const double *lut_v, int64_t frame, int i, rr_sim_streak *res);

__global__ void k_sim_vt(const double *lut_d, double *lut_v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < SIM_NLUT) lut_v[i] = sim_v_terminal(lut_d[i]);
}

__global__ void k_sim_frame(sim_consts c, const double *cdf, const double *lut_d, const double *lut_v, int64_t frame, int n_cand, int cap,
                            rr_sim_streak *out, int *counter) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cand) return;
    rr_sim_streak r;
    if (!sim_candidate(c, cdf, lut_d, lut_v, frame, i, &r)) return;
    int slot = atomicAdd(counter, 1);
    if (slot >= cap) return;
    out[slot] = r;
}

__device__ bool sim_candidate(const sim_consts &c, const double *cdf, const double *lut_d, const double *lut_v, int64_t frame, int i, rr_sim_streak *res) {
    // diameter: inverse transform on the LUT (the binary scans a 50001-entry table linearly)
    double u = sim_u01(c.seed, (uint64_t)frame, i, 0);
    int lo = 0, hi = SIM_NLUT - 1;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (cdf[mid] < u) lo = mid + 1; else hi = mid; }
    double D = lut_d[lo];
    double v = lut_v[lo];                 // sim_v_terminal(D), tabulated
    double zmax;
    sim_region(c, D, v, &zmax);
    if (zmax <= c.z0) return false;
    // position uniform in the region: depth with density proportional to the cross-section area
    double a = (c.W + 2 * c.margin_x) * c.H / (c.f_px * c.f_px), b = (c.W + 2 * c.margin_x) / c.f_px * (v * c.T);
    double F0 = a * c.z0 * c.z0 * c.z0 / 3 + b * c.z0 * c.z0 / 2, F1 = a * zmax * zmax * zmax / 3 + b * zmax * zmax / 2;
    double target = F0 + sim_u01(c.seed, (uint64_t)frame, i, 1) * (F1 - F0);
    double zl = c.z0, zh = zmax;
    for (int it = 0; it < 60; it++) { double zm = 0.5 * (zl + zh); if (a * zm * zm * zm / 3 + b * zm * zm / 2 < target) zl = zm; else zh = zm; }
    double z = 0.5 * (zl + zh);
    double half_w = (c.W / 2 + c.margin_x) * z / c.f_px;
    double x = (2 * sim_u01(c.seed, (uint64_t)frame, i, 2) - 1) * half_w;
    double ybot = -(c.H / 2) * z / c.f_px, ytop = (c.H / 2) * z / c.f_px + v * c.T;
    double y = ybot + sim_u01(c.seed, (uint64_t)frame, i, 3) * (ytop - ybot);
    // shutter open snapshot
    double x1 = x, y1 = y, z1 = z;
    // the binary's stepper over the exposure: semi-implicit Euler, vertical force only (dt is float32 there)
    double vy = -v;
    bool constant = false;
    const double m = sim_mass(D);
    for (int s = 0; s < c.n_steps; s++) {
        if (!constant) {
            double acc = -(m * 9.81 - sim_drag(D, -vy)) / m;
            vy += acc * c.dt;
            if (acc >= -0.1) constant = true;
        }
        y += vy * c.dt;
        z -= c.vcam * c.dt;             // camera moves forward: relative drift towards the camera
    }
    double x2 = x, y2 = y, z2 = z;
    if (z2 <= 1e-3) return false;
    double u1 = c.W / 2 + c.f_px * x1 / z1, v1 = c.H / 2 + c.f_px * y1 / z1;     // y up
    double u2 = c.W / 2 + c.f_px * x2 / z2, v2 = c.H / 2 + c.f_px * y2 / z2;
    bool in1 = u1 >= 0 && u1 < c.W && v1 >= 0 && v1 < c.H, in2 = u2 >= 0 && u2 < c.W && v2 >= 0 && v2 < c.H;
    if (!(in1 || in2)) return false;                                            // IsIn()
    double w1 = D * c.f_px / z1, w2 = D * c.f_px / z2;
    if ((w1 > w2 ? w1 : w2) < c.wmin) return false;                             // IsFoglike()
    rr_sim_streak r;
    r.wp1[0] = x1; r.wp1[1] = y1; r.wp1[2] = -z1;
    r.wp2[0] = x2; r.wp2[1] = y2; r.wp2[2] = -z2;
    r.wd1 = D; r.wd2 = D;
    r.ip1[0] = u1; r.ip1[1] = v1; r.ip2[0] = u2; r.ip2[1] = v2;
    r.iw1 = w1; r.iw2 = w2;
    r.pid = i;
    *res = r;
    return true;
}

struct rr_context;
extern "C" int rr_sim_device_of(rr_context *c);   // rr_api.cu
extern "C" void rr_set_error(const char *msg);
extern "C" void *rr_ctx_scratch(rr_context *c, int which, size_t bytes);      // rr_api.cu: grow-only device buffers of the context
extern "C" void *rr_ctx_stream(rr_context *c);
extern "C" void rr_ctx_count_launches(rr_context *c, int n);

// which parameter set the tables in the context's scratch buffer 0 (diameter CDF, diameters, terminal velocities) belong to
static thread_local rr_sim_params g_tab_params;
static thread_local sim_consts g_tab_consts;
static thread_local double g_tab_mean = 0;
static thread_local rr_context *g_tab_ctx = nullptr;

// ---- host-side preparation shared by the two entry points ---------------------------------------------------------------
static int sim_prepare(const rr_sim_params *p, sim_consts *cc, std::vector<double> *cdf_out, std::vector<double> *d_out, double *mean_out) {
    if (p->W <= 0 || p->H <= 0 || p->focal_m <= 0 || p->pix_size_m <= 0 || p->fallrate_mmh <= 0 || p->sim_hz <= 0 || p->z_far <= p->z_near ||
        p->d_max_mm <= p->d_min_mm || p->min_width_px <= 0) { rr_set_error("rr_simulate_particles: invalid parameters"); return RR_ERR_ARG; }
    sim_consts c;
    c.f_px = p->focal_m / p->pix_size_m; c.W = p->W; c.H = p->H; c.T = p->exposure_ms / 1000.0;
    c.dt = (double)(float)(1.0 / p->sim_hz);
    c.n_steps = (int)lrint(c.T * p->sim_hz); if (c.n_steps < 1) c.n_steps = 1;
    c.dt = c.T / c.n_steps < c.dt ? c.T / c.n_steps : c.dt;     // exposure shorter than a step: one partial step
    if (fabs(c.n_steps * c.dt - c.T) > 1e-12) c.dt = c.T / c.n_steps;
    c.z0 = p->z_near; c.z1 = p->z_far; c.wmin = p->min_width_px; c.vcam = p->cam_speed_kmh * 1000.0 / 3600.0;
    c.dmin = p->d_min_mm * 1e-3; c.dmax = p->d_max_mm * 1e-3; c.margin_x = 2.0; c.seed = p->seed;
    // LUT over diameters: airborne concentration c(D) = K N(D) / v(D)  (emission pdf ~ N(D), budget R),
    // weighted by the visible-region volume; mean count = sum
    const double lam = 4.1 * pow(p->fallrate_mmh, -0.21);                 // per mm
    std::vector<double> d(SIM_NLUT), w(SIM_NLUT), cdf(SIM_NLUT);
    double water = 0;                                                     // integral N(D) V(D) dD  [m^3 / m^3]
    const double dD_mm = (p->d_max_mm - p->d_min_mm) / SIM_NLUT;
    for (int i = 0; i < SIM_NLUT; i++) {
        double Dmm = p->d_min_mm + (i + 0.5) * dD_mm;
        d[i] = Dmm * 1e-3;
        double N = 8000.0 * exp(-lam * Dmm);                              // m^-3 mm^-1
        water += N * (SIM_PI / 6) * pow(d[i], 3) * dD_mm;
    }
    const double K = (p->fallrate_mmh * 1e-3 / 3600.0) / water;           // m/s
    double mean = 0;
    for (int i = 0; i < SIM_NLUT; i++) {
        double Dmm = d[i] * 1e3, N = 8000.0 * exp(-lam * Dmm);
        double v = sim_v_terminal(d[i]), zmax;
        double vol = sim_region(c, d[i], v, &zmax);
        w[i] = K * N / v * dD_mm * vol;
        mean += w[i];
    }
    double acc = 0;
    for (int i = 0; i < SIM_NLUT; i++) { acc += w[i]; cdf[i] = mean > 0 ? acc / mean : 1.0; }
    cdf[SIM_NLUT - 1] = 1.0;
    *cc = c; *cdf_out = cdf; *d_out = d; *mean_out = mean;
    return RR_OK;
}

// candidates of one frame: Poisson count (normal approximation above 64, Knuth below), counter-based
static int sim_poisson(const sim_consts &c, double mean, int64_t frame) {
    int n_cand;
    if (mean > 64) {
        double u1 = sim_u01(c.seed, (uint64_t)frame, 0xFFFFFFFFull, 7), u2 = sim_u01(c.seed, (uint64_t)frame, 0xFFFFFFFFull, 8);
        double g = sqrt(-2 * log(u1)) * cos(2 * SIM_PI * u2);
        n_cand = (int)lrint(mean + sqrt(mean) * g);
    } else {
        double L = exp(-mean), pacc = 1; n_cand = -1; uint64_t k = 0;
        do { n_cand++; pacc *= sim_u01(c.seed, (uint64_t)frame, 0xFFFFFFFFull, 9 + k++); } while (pacc > L);
    }
    return n_cand < 0 ? 0 : n_cand;
}

extern "C" int rr_simulate_particles(rr_context *ctx, const rr_sim_params *p, int64_t first_frame, int n_frames, int max_per_frame,
                                     rr_sim_streak *out, int32_t *counts, double *expected_per_frame) {
    if (!ctx || !p || !out || !counts || n_frames <= 0 || max_per_frame <= 0) { rr_set_error("rr_simulate_particles: bad arguments"); return RR_ERR_ARG; }
    if (cudaSetDevice(rr_sim_device_of(ctx)) != cudaSuccess) { rr_set_error("rr_simulate_particles: cudaSetDevice failed"); return RR_ERR_CUDA; }
    sim_consts c;
    std::vector<double> cdf, d;
    double mean = 0;
    int rc = sim_prepare(p, &c, &cdf, &d, &mean);
    if (rc != RR_OK) return rc;
    if (expected_per_frame) *expected_per_frame = mean;
    // tables and output slices live in the context's grow-only scratch buffers; frames go to the device in chunks: all
    // launches of a chunk first, one read-back of its counters, then exactly the records that were produced
    cudaStream_t st = (cudaStream_t)rr_ctx_stream(ctx);
    const int CH = n_frames < 64 ? n_frames : 64;
    double *d_tab = (double *)rr_ctx_scratch(ctx, 0, sizeof(double) * 3 * SIM_NLUT);
    int *d_cnt = (int *)rr_ctx_scratch(ctx, 4, sizeof(int) * 64 + 1024);
    rr_sim_streak *d_out = (rr_sim_streak *)rr_ctx_scratch(ctx, 1, sizeof(rr_sim_streak) * (size_t)CH * max_per_frame);
    if (!d_tab || !d_cnt || !d_out) { rr_set_error("rr_simulate_particles: out of device memory"); return RR_ERR_CUDA; }
    double *d_cdf = d_tab, *d_d = d_tab + SIM_NLUT, *d_v = d_tab + 2 * SIM_NLUT;
    cudaError_t e;
#define SIMCK(call) if ((e = (call)) != cudaSuccess) { rr_set_error(cudaGetErrorString(e)); return RR_ERR_CUDA; }
    SIMCK(cudaMemcpyAsync(d_cdf, cdf.data(), sizeof(double) * SIM_NLUT, cudaMemcpyHostToDevice, st));
    SIMCK(cudaMemcpyAsync(d_d, d.data(), sizeof(double) * SIM_NLUT, cudaMemcpyHostToDevice, st));
    k_sim_vt<<<SIM_NLUT / 128, 128, 0, st>>>(d_d, d_v);
    SIMCK(cudaStreamSynchronize(st));
    { rr_sim_params key = *p; key.seed = 0; g_tab_params = key; g_tab_consts = c; g_tab_mean = mean; g_tab_ctx = ctx; }   // what scratch buffer 0 now holds
    std::vector<int> cnt(CH);
    for (int f0 = 0; f0 < n_frames && rc == RR_OK; f0 += CH) {
        const int nf = n_frames - f0 < CH ? n_frames - f0 : CH;
        SIMCK(cudaMemsetAsync(d_cnt, 0, sizeof(int) * nf, st));
        for (int f = 0; f < nf; f++) {
            const int64_t frame = first_frame + f0 + f;
            const int n_cand = sim_poisson(c, mean, frame);
            if (n_cand > 0) k_sim_frame<<<(n_cand + 127) / 128, 128, 0, st>>>(c, d_cdf, d_d, d_v, frame, n_cand, max_per_frame, d_out + (size_t)f * max_per_frame, d_cnt + f);
        }
        SIMCK(cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(int) * nf, cudaMemcpyDeviceToHost, st));
        SIMCK(cudaStreamSynchronize(st));
        for (int f = 0; f < nf; f++) {
            int n = cnt[f];
            if (n > max_per_frame) { rr_set_error("rr_simulate_particles: more streaks than max_per_frame"); rc = RR_ERR_CAPACITY; n = max_per_frame; }
            counts[f0 + f] = n;
            if (n) SIMCK(cudaMemcpyAsync(out + (size_t)(f0 + f) * max_per_frame, d_out + (size_t)f * max_per_frame, sizeof(rr_sim_streak) * n, cudaMemcpyDeviceToHost, st));
        }
        SIMCK(cudaStreamSynchronize(st));
    }
#undef SIMCK
    return rc;
}

// ---- device-resident form: simulator -> loader arithmetic -> in-frame filter -> RNG draws, records stay in HBM ----------
// One launch simulates every candidate of every frame and turns the imaged ones into rr_streak_rec exactly as
// DBManager.load_streaks_from_xml would from the simulator's XML (common/bad_weather.py:200-238; same float64 operations in
// the same order as rain_rendering_b200/streaks.py: records_from_raw) and applies the in-frame filter (generator.py:413-420);
// a block per frame then compacts the survivors in candidate (= pid) order and one thread walks them with the frame's NumPy
// legacy MT19937 stream (np.random.seed(frame index), generator.py:318): randint for the texture (bad_weather.py:252-264),
// normal for the wind noise of non-Big drops (generator.py:136; with noise_std == 0 only its draws are consumed).
struct sim_mt {                       // the NumPy legacy stream (csrc/rr_host.cpp holds the host twin)
    uint32_t *key;
    int pos, has_gauss;
    double gauss;
    __device__ void seed(uint32_t s) {
        for (int i = 0; i < 624; i++) { key[i] = s; s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)i + 1u; }
        pos = 624; has_gauss = 0; gauss = 0.0;
    }
    __device__ void gen() {
        const uint32_t N = 624, M = 397, A = 0x9908b0dfu, UP = 0x80000000u, LO = 0x7fffffffu;
        uint32_t y, i;
        for (i = 0; i < N - M; i++) { y = (key[i] & UP) | (key[i + 1] & LO); key[i] = key[i + M] ^ (y >> 1) ^ ((0u - (y & 1u)) & A); }
        for (; i < N - 1; i++) { y = (key[i] & UP) | (key[i + 1] & LO); key[i] = key[i + M - N] ^ (y >> 1) ^ ((0u - (y & 1u)) & A); }
        y = (key[N - 1] & UP) | (key[0] & LO);
        key[N - 1] = key[M - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & A);
        pos = 0;
    }
    __device__ uint32_t next32() {
        if (pos == 624) gen();
        uint32_t y = key[pos++];
        y ^= (y >> 11); y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= (y >> 18);
        return y;
    }
    __device__ double next_double() {
        const int32_t a = (int32_t)(next32() >> 5), b = (int32_t)(next32() >> 6);
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
    __device__ uint32_t bounded9() {              // legacy randint(low, low + 10): range 9, mask 15
        uint32_t v;
        do { v = next32() & 15u; } while (v > 9u);
        return v;
    }
    __device__ double legacy_gauss() {
        if (has_gauss) { const double t = gauss; has_gauss = 0; gauss = 0.0; return t; }
        double f, x1, x2, r2;
        do {
            x1 = 2.0 * next_double() - 1.0;
            x2 = 2.0 * next_double() - 1.0;
            r2 = x1 * x1 + x2 * x2;
        } while (r2 >= 1.0 || r2 == 0.0);
        f = sqrt(-2.0 * log(r2) / r2);
        gauss = f * x1; has_gauss = 1;
        return f * x2;
    }
};

__global__ void __launch_bounds__(128) k_sim_records(sim_consts c, const double *cdf, const double *lut_d, const double *lut_v, int64_t first_frame, const int *n_cand,
                                                    int max_cand, int render_scale, int W, int H, rr_streak_rec *cand, unsigned char *flags) {
    const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= max_cand) return;
    unsigned char keep = 0;
    rr_sim_streak s;
    if (i < n_cand[f] && sim_candidate(c, cdf, lut_d, lut_v, first_frame + f, i, &s)) {
        // DBManager.load_streaks_from_xml (bad_weather.py:208-238) on the simulator's values
        const double rs = (double)render_scale;
        double p1x = s.ip1[0] / rs, p1y = s.ip1[1] / rs, p2x = s.ip2[0] / rs, p2y = s.ip2[1] / rs;       // :208-209
        const double iw1 = s.iw1 / rs, iw2 = s.iw2 / rs;                                                   // :210-211
        p1y = (double)H - p1y; p2y = (double)H - p2y;                                                      // :221-222
        rr_streak_rec r;
        r.wp1[0] = s.wp1[0]; r.wp1[1] = s.wp1[1]; r.wp1[2] = s.wp1[2] * -1;                                // :223-224
        r.wp2[0] = s.wp2[0]; r.wp2[1] = s.wp2[1]; r.wp2[2] = s.wp2[2] * -1;
        const double dx = fabs(p1x - p2x), dy = fabs(p1y - p2y);
        const long long max_width = (long long)(iw1 > iw2 ? iw1 : iw2);                                    // :226 int() truncation
        const double nrm = sqrt(fma(dy, dy, dx * dx));                                                     // np.linalg.norm, :229
        const double cos_theta = 0.0 * (dx / nrm) + -1.0 * (-(dy / nrm));                                  // :228-231
        r.ratio = (double)max_width / (dy / cos_theta);                                                    // :232-233
        const long long x1 = (long long)rint(p1x), y1 = (long long)rint(p1y), x2 = (long long)rint(p2x), y2 = (long long)rint(p2y);   // :234-235 half-even
        const double ddx = (double)(x1 - x2), ddy = (double)(y1 - y2);
        const long long length = (long long)ceil(sqrt(ddx * ddx + ddy * ddy));                             // :236
        r.iw1 = iw1; r.iw2 = iw2; r.noise_deg = 0.0;
        r.ip1[0] = r.ip1m[0] = (int32_t)x1; r.ip1[1] = r.ip1m[1] = (int32_t)y1;
        r.ip2[0] = r.ip2m[0] = (int32_t)x2; r.ip2[1] = r.ip2m[1] = (int32_t)y2;
        r.max_width = (int32_t)max_width; r.length = (int32_t)length; r.pid = (int32_t)s.pid;
        r.type = max_width >= 4 ? 0 : (max_width > 1 ? 1 : 2);                                             // :99-106
        r.tex_idx = 0; r.pad[0] = r.pad[1] = 0;
        const int m = H > W ? H : W;
        const bool loaded = max_width >= 1 && length >= 1;                                                 // :238
        const bool ok_w = 1 <= max_width && max_width < m, ok_l = 1 <= length && length < m;               // generator.py:413-420
        const bool ins = 0 <= x1 && x1 < W && 0 <= y1 && y1 < H, ine = 0 <= x2 && x2 < W && 0 <= y2 && y2 < H;
        if (loaded && ok_w && ok_l && (ins || ine)) { keep = 1; cand[(size_t)f * max_cand + i] = r; }
    }
    flags[(size_t)f * max_cand + i] = keep;
}

__global__ void __launch_bounds__(256) k_sim_count(const unsigned char *flags, int max_cand, int *counts) {
    const int f = blockIdx.x;
    int n = 0;
    for (int i = threadIdx.x; i < max_cand; i += 256) n += flags[(size_t)f * max_cand + i];
    __shared__ int sh[8];
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int k = 0; k < 8; k++) t += sh[k]; counts[f] = t; }
}

__global__ void k_sim_offsets(const int *counts, int F, int32_t *offsets) {
    if (blockIdx.x == 0 && threadIdx.x == 0) { int o = 0; offsets[0] = 0; for (int f = 0; f < F; f++) { o += counts[f]; offsets[f + 1] = o; } }
}

__global__ void __launch_bounds__(256) k_sim_compact_draw(const rr_streak_rec *cand, const unsigned char *flags, int max_cand, const int32_t *offsets,
                                                         int64_t first_frame, double r0, double r1, double r2, double r3, int n_ratios,
                                                         double noise_std, double noise_scale, rr_streak_rec *out) {
    __shared__ uint32_t key[624];
    __shared__ int wsum[8];
    __shared__ int base;
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) base = offsets[f];
    __syncthreads();
    // ordered compaction: candidate order = pid order = the order the loader's dict and the in-frame filter keep
    for (int i0 = 0; i0 < max_cand; i0 += 256) {
        const int i = i0 + tid;
        const int k = i < max_cand ? flags[(size_t)f * max_cand + i] : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, k);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        int pre = base;
        for (int w = 0; w < warp; w++) pre += wsum[w];
        if (k) {
            const int4 *src = (const int4 *)(cand + (size_t)f * max_cand + i);
            int4 *dst = (int4 *)(out + pre + __popc(bal & ((1u << lane) - 1)));
#pragma unroll
            for (int q = 0; q < (int)(sizeof(rr_streak_rec) / sizeof(int4)); q++) dst[q] = src[q];
        }
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < 8; w++) t += wsum[w]; base += t; }
        __syncthreads();
    }
    if (tid != 0) return;
    __threadfence_block();
    sim_mt mt;
    mt.key = key;
    mt.seed((uint32_t)(first_frame + f));                          // np.random.seed(frame index), generator.py:318
    const double ratios[4] = {r0, r1, r2, r3};
    const int nr = n_ratios < 4 ? n_ratios : 4;
    for (int s = offsets[f]; s < offsets[f + 1]; s++) {
        rr_streak_rec &r = out[s];
        int b = 0;
        while (b < nr && !(r.ratio < ratios[b])) b++;
        r.tex_idx = (uint8_t)(10 * b + (int)mt.bounded9());        // bad_weather.py:252-264
        r.noise_deg = r.type != 0 ? (0.0 + noise_std * mt.legacy_gauss()) * noise_scale : 0.0;      // generator.py:136
    }
}


extern "C" int rr_simulate_records_device(rr_context *ctx, const rr_sim_params *p, int64_t first_frame, int n_frames, int render_scale,
                                          const double *db_ratios, int n_ratios, double noise_std, double noise_scale,
                                          rr_streak_rec **d_records, int32_t *h_offsets, double *expected_per_frame) {
    if (!ctx || !p || !d_records || !h_offsets || n_frames <= 0 || render_scale < 1 || (n_ratios > 0 && !db_ratios)) { rr_set_error("rr_simulate_records_device: bad arguments"); return RR_ERR_ARG; }
    if (noise_std != 0.0 && noise_scale != 0.0) {
        rr_set_error("rr_simulate_records_device: wind noise couples consecutive frames through the write-back of generator.py:152-161; "
                     "use rr_simulate_particles + the host record assembly for noise_std * noise_scale != 0");
        return RR_ERR_ARG;
    }
    if (first_frame < 0 || first_frame + n_frames > 0xffffffffll) { rr_set_error("rr_simulate_records_device: frame index out of the 32-bit seed range"); return RR_ERR_ARG; }
    if (cudaSetDevice(rr_sim_device_of(ctx)) != cudaSuccess) { rr_set_error("rr_simulate_records_device: cudaSetDevice failed"); return RR_ERR_CUDA; }
    // the tables of a parameter set (diameter CDF, diameters, terminal velocities) are built once and stay on the device
    rr_sim_params key = *p;
    key.seed = 0;
    const bool hit = g_tab_ctx == ctx && memcmp(&key, &g_tab_params, sizeof(key)) == 0 && rr_ctx_scratch(ctx, 0, 0) != nullptr;
    sim_consts c;
    std::vector<double> cdf, d;
    double mean = 0;
    int rc = RR_OK;
    if (hit) { c = g_tab_consts; c.seed = p->seed; mean = g_tab_mean; }
    else {
        rc = sim_prepare(p, &c, &cdf, &d, &mean);
        if (rc != RR_OK) return rc;
    }
    if (expected_per_frame) *expected_per_frame = mean;
    std::vector<int> n_cand(n_frames);
    int max_cand = 1;
    for (int f = 0; f < n_frames; f++) { n_cand[f] = sim_poisson(c, mean, first_frame + f); if (n_cand[f] > max_cand) max_cand = n_cand[f]; }
    const int W = p->W / render_scale, H = p->H / render_scale;
    cudaStream_t st = (cudaStream_t)rr_ctx_stream(ctx);
    const size_t F = (size_t)n_frames;
    double *d_tab = (double *)rr_ctx_scratch(ctx, 0, sizeof(double) * 3 * SIM_NLUT);
    int *d_ints = (int *)rr_ctx_scratch(ctx, 4, sizeof(int) * (2 * F + 2) + sizeof(int32_t) * (F + 1));
    rr_streak_rec *d_cand = (rr_streak_rec *)rr_ctx_scratch(ctx, 1, sizeof(rr_streak_rec) * F * max_cand);
    unsigned char *d_flags = (unsigned char *)rr_ctx_scratch(ctx, 2, F * max_cand);
    rr_streak_rec *d_out = (rr_streak_rec *)rr_ctx_scratch(ctx, 3, sizeof(rr_streak_rec) * F * max_cand);
    if (!d_tab || !d_ints || !d_cand || !d_flags || !d_out) { rr_set_error("rr_simulate_records_device: out of device memory"); g_tab_ctx = nullptr; return RR_ERR_CUDA; }
    double *d_cdf = d_tab, *d_d = d_tab + SIM_NLUT, *d_v = d_tab + 2 * SIM_NLUT;
    int *d_ncand = d_ints, *d_counts = d_ncand + F;
    int32_t *d_offsets = (int32_t *)(d_counts + F + 2);
    cudaError_t e;
#define SIMCK(call) if ((e = (call)) != cudaSuccess) { rr_set_error(cudaGetErrorString(e)); g_tab_ctx = nullptr; return RR_ERR_CUDA; }
    if (!hit) {
        SIMCK(cudaMemcpyAsync(d_cdf, cdf.data(), sizeof(double) * SIM_NLUT, cudaMemcpyHostToDevice, st));
        SIMCK(cudaMemcpyAsync(d_d, d.data(), sizeof(double) * SIM_NLUT, cudaMemcpyHostToDevice, st));
        k_sim_vt<<<SIM_NLUT / 128, 128, 0, st>>>(d_d, d_v);
        SIMCK(cudaStreamSynchronize(st));           // cdf / d are host vectors about to go out of scope
        g_tab_params = key; g_tab_consts = c; g_tab_mean = mean; g_tab_ctx = ctx;
        rr_ctx_count_launches(ctx, 1);
    }
    SIMCK(cudaMemcpyAsync(d_ncand, n_cand.data(), sizeof(int) * F, cudaMemcpyHostToDevice, st));
    dim3 g((max_cand + 127) / 128, n_frames);
    k_sim_records<<<g, 128, 0, st>>>(c, d_cdf, d_d, d_v, first_frame, d_ncand, max_cand, render_scale, W, H, d_cand, d_flags);
    k_sim_count<<<n_frames, 256, 0, st>>>(d_flags, max_cand, d_counts);
    k_sim_offsets<<<1, 32, 0, st>>>(d_counts, n_frames, d_offsets);
    const double rr[4] = {n_ratios > 0 ? db_ratios[0] : 0, n_ratios > 1 ? db_ratios[1] : 0, n_ratios > 2 ? db_ratios[2] : 0, n_ratios > 3 ? db_ratios[3] : 0};
    k_sim_compact_draw<<<n_frames, 256, 0, st>>>(d_cand, d_flags, max_cand, d_offsets, first_frame, rr[0], rr[1], rr[2], rr[3], n_ratios, noise_std,
                                                  noise_scale, d_out);
    SIMCK(cudaGetLastError());
    rr_ctx_count_launches(ctx, 4);
    // the only thing that comes back: the n + 1 frame offsets (the render entry points take them as a host array)
    SIMCK(cudaMemcpyAsync(h_offsets, d_offsets, sizeof(int32_t) * (F + 1), cudaMemcpyDeviceToHost, st));
    SIMCK(cudaStreamSynchronize(st));
#undef SIMCK
    *d_records = d_out;
    return RR_OK;
}

// host-side evaluation of the force model for the CPU tests (no GPU needed)
extern "C" void rr_host_sim_physics(double D_m, double *v_terminal, double *drag_at_vt, double *mass) {
    double v = sim_v_terminal(D_m);
    if (v_terminal) *v_terminal = v;
    if (drag_at_vt) *drag_at_vt = sim_drag(D_m, v);
    if (mass) *mass = sim_mass(D_m);
}
