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

__global__ void k_sim_frame(sim_consts c, const double *cdf, const double *lut_d, int64_t frame, int n_cand, int cap,
                            rr_sim_streak *out, int *counter) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cand) return;
    // diameter: inverse transform on the LUT (the binary scans a 50001-entry table linearly)
    double u = sim_u01(c.seed, (uint64_t)frame, i, 0);
    int lo = 0, hi = SIM_NLUT - 1;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (cdf[mid] < u) lo = mid + 1; else hi = mid; }
    double D = lut_d[lo];
    double v = sim_v_terminal(D);
    double zmax;
    sim_region(c, D, v, &zmax);
    if (zmax <= c.z0) return;
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
    if (z2 <= 1e-3) return;
    double u1 = c.W / 2 + c.f_px * x1 / z1, v1 = c.H / 2 + c.f_px * y1 / z1;     // y up
    double u2 = c.W / 2 + c.f_px * x2 / z2, v2 = c.H / 2 + c.f_px * y2 / z2;
    bool in1 = u1 >= 0 && u1 < c.W && v1 >= 0 && v1 < c.H, in2 = u2 >= 0 && u2 < c.W && v2 >= 0 && v2 < c.H;
    if (!(in1 || in2)) return;                                                  // IsIn()
    double w1 = D * c.f_px / z1, w2 = D * c.f_px / z2;
    if ((w1 > w2 ? w1 : w2) < c.wmin) return;                                   // IsFoglike()
    int slot = atomicAdd(counter, 1);
    if (slot >= cap) return;
    rr_sim_streak r;
    r.wp1[0] = x1; r.wp1[1] = y1; r.wp1[2] = -z1;
    r.wp2[0] = x2; r.wp2[1] = y2; r.wp2[2] = -z2;
    r.wd1 = D; r.wd2 = D;
    r.ip1[0] = u1; r.ip1[1] = v1; r.ip2[0] = u2; r.ip2[1] = v2;
    r.iw1 = w1; r.iw2 = w2;
    r.pid = i;
    out[slot] = r;
}

struct rr_context;
extern "C" int rr_sim_device_of(rr_context *c);   // rr_api.cu
extern "C" void rr_set_error(const char *msg);

extern "C" int rr_simulate_particles(rr_context *ctx, const rr_sim_params *p, int64_t first_frame, int n_frames, int max_per_frame,
                                     rr_sim_streak *out, int32_t *counts, double *expected_per_frame) {
    if (!ctx || !p || !out || !counts || n_frames <= 0 || max_per_frame <= 0) { rr_set_error("rr_simulate_particles: bad arguments"); return RR_ERR_ARG; }
    if (p->W <= 0 || p->H <= 0 || p->focal_m <= 0 || p->pix_size_m <= 0 || p->fallrate_mmh <= 0 || p->sim_hz <= 0 || p->z_far <= p->z_near ||
        p->d_max_mm <= p->d_min_mm || p->min_width_px <= 0) { rr_set_error("rr_simulate_particles: invalid parameters"); return RR_ERR_ARG; }
    if (cudaSetDevice(rr_sim_device_of(ctx)) != cudaSuccess) { rr_set_error("rr_simulate_particles: cudaSetDevice failed"); return RR_ERR_CUDA; }
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
    if (expected_per_frame) *expected_per_frame = mean;
    double acc = 0;
    for (int i = 0; i < SIM_NLUT; i++) { acc += w[i]; cdf[i] = mean > 0 ? acc / mean : 1.0; }
    cdf[SIM_NLUT - 1] = 1.0;
    double *d_cdf = nullptr, *d_d = nullptr; rr_sim_streak *d_out = nullptr; int *d_cnt = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc(&d_cdf, sizeof(double) * SIM_NLUT)) != cudaSuccess || (e = cudaMalloc(&d_d, sizeof(double) * SIM_NLUT)) != cudaSuccess ||
        (e = cudaMalloc(&d_out, sizeof(rr_sim_streak) * (size_t)max_per_frame)) != cudaSuccess || (e = cudaMalloc(&d_cnt, sizeof(int))) != cudaSuccess) {
        rr_set_error(cudaGetErrorString(e)); cudaFree(d_cdf); cudaFree(d_d); cudaFree(d_out); cudaFree(d_cnt); return RR_ERR_CUDA;
    }
    cudaMemcpy(d_cdf, cdf.data(), sizeof(double) * SIM_NLUT, cudaMemcpyHostToDevice);
    cudaMemcpy(d_d, d.data(), sizeof(double) * SIM_NLUT, cudaMemcpyHostToDevice);
    int rc = RR_OK;
    for (int f = 0; f < n_frames && rc == RR_OK; f++) {
        int64_t frame = first_frame + f;
        // Poisson count (normal approximation above 64, Knuth below), counter-based
        int n_cand;
        if (mean > 64) {
            double u1 = sim_u01(c.seed, (uint64_t)frame, 0xFFFFFFFFull, 7), u2 = sim_u01(c.seed, (uint64_t)frame, 0xFFFFFFFFull, 8);
            double g = sqrt(-2 * log(u1)) * cos(2 * SIM_PI * u2);
            n_cand = (int)lrint(mean + sqrt(mean) * g);
        } else {
            double L = exp(-mean), pacc = 1; n_cand = -1; uint64_t k = 0;
            do { n_cand++; pacc *= sim_u01(c.seed, (uint64_t)frame, 0xFFFFFFFFull, 9 + k++); } while (pacc > L);
        }
        if (n_cand < 0) n_cand = 0;
        cudaMemset(d_cnt, 0, sizeof(int));
        if (n_cand > 0) k_sim_frame<<<(n_cand + 127) / 128, 128>>>(c, d_cdf, d_d, frame, n_cand, max_per_frame, d_out, d_cnt);
        int cnt = 0;
        if ((e = cudaMemcpy(&cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost)) != cudaSuccess) { rr_set_error(cudaGetErrorString(e)); rc = RR_ERR_CUDA; break; }
        if (cnt > max_per_frame) { rr_set_error("rr_simulate_particles: more streaks than max_per_frame"); rc = RR_ERR_CAPACITY; cnt = max_per_frame; }
        counts[f] = cnt;
        if (cnt) cudaMemcpy(out + (size_t)f * max_per_frame, d_out, sizeof(rr_sim_streak) * cnt, cudaMemcpyDeviceToHost);
    }
    cudaFree(d_cdf); cudaFree(d_d); cudaFree(d_out); cudaFree(d_cnt);
    return rc;
}

// host-side evaluation of the force model for the CPU tests (no GPU needed)
extern "C" void rr_host_sim_physics(double D_m, double *v_terminal, double *drag_at_vt, double *mass) {
    double v = sim_v_terminal(D_m);
    if (v_terminal) *v_terminal = v;
    if (drag_at_vt) *drag_at_vt = sim_drag(D_m, v);
    if (mass) *mass = sim_mass(D_m);
}
