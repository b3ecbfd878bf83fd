// Host-side helpers exported through the C ABI (no CUDA): a bit-exact mirror of the two NumPy
// legacy-RNG draws the reference makes per streak, so the per-frame path needs no Python-level
// RNG calls.
//   np.random.seed(frame_idx)                      common/generator.py:318
//   np.random.randint(10*b, 10*b + 10)             common/bad_weather.py:252-264   (every streak)
//   np.random.normal(0.0, noise_std) * noise_scale common/generator.py:136         (non-Big streaks)
// NumPy legacy stream (numpy/random/_mt19937.pyx, src/legacy/legacy-distributions.c):
// MT19937 seeded by init_genrand; randint = masked rejection on 32-bit draws; normal = polar
// Box-Muller with a cached second variate on 53-bit doubles.
#include <math.h>
#include <stdint.h>
#include "../../include/rain_b200.h"
#include "rr_streak_geom.h"

namespace {
struct MT {
    uint32_t key[624];
    int pos;
    int has_gauss;
    double gauss;
    void seed(uint32_t s) {
        for (int i = 0; i < 624; i++) {
            key[i] = s;
            s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)i + 1u;
        }
        pos = 624; has_gauss = 0; gauss = 0.0;
    }
    void gen() {
        const uint32_t N = 624, M = 397, MATRIX_A = 0x9908b0dfu, UPPER = 0x80000000u, LOWER = 0x7fffffffu;
        uint32_t y;
        uint32_t i;
        for (i = 0; i < N - M; i++) { y = (key[i] & UPPER) | (key[i + 1] & LOWER); key[i] = key[i + M] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A); }
        for (; i < N - 1; i++) { y = (key[i] & UPPER) | (key[i + 1] & LOWER); key[i] = key[i + (M - N)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A); }
        y = (key[N - 1] & UPPER) | (key[0] & LOWER);
        key[N - 1] = key[M - 1] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
        pos = 0;
    }
    uint32_t next32() {
        if (pos == 624) gen();
        uint32_t y = key[pos++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    double next_double() {
        int32_t a = next32() >> 5, b = next32() >> 6;
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
    // legacy randint(low, low + 10): rng = 9, mask = 15
    uint32_t bounded(uint32_t rng) {
        uint32_t mask = rng;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        uint32_t v;
        do { v = next32() & mask; } while (v > rng);
        return v;
    }
    double legacy_gauss() {
        if (has_gauss) {
            double t = gauss;
            has_gauss = 0; gauss = 0.0;
            return t;
        }
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
}  // namespace

extern "C" int rr_host_draw_randoms(uint32_t seed, int n, const uint8_t *types, const int32_t *buckets, double noise_std,
                                    double noise_scale, uint8_t *tex_idx, double *noise_deg) {
    if (n < 0 || (n > 0 && (!types || !buckets || !tex_idx || !noise_deg))) return RR_ERR_ARG;
    MT mt;
    mt.seed(seed);
    for (int i = 0; i < n; i++) {
        int lo = 10 * buckets[i];
        tex_idx[i] = (uint8_t)(lo + (int)mt.bounded(9));
        if (types[i] != 0) noise_deg[i] = (0.0 + noise_std * mt.legacy_gauss()) * noise_scale;
        else noise_deg[i] = 0.0;
    }
    return RR_OK;
}

// A batch of image frames' records from their simulator frames, one native call: the in-frame filter
// (common/generator.py:413-420, on the current -- possibly wind-mutated -- positions), the texture bucket
// (common/bad_weather.py:250-265: first i with ratio < ratios[i], else 4) and the frame's two NumPy RNG draws per
// streak (generator.py:318,136; bad_weather.py:252-264).  Output records carry ip1 = ip1m, ip2 = ip2m (the angle is taken
// from the positions before this frame's wind rotation, generator.py:138-144).  The rotation itself (generator.py:149-161)
// is NOT applied here: with noise_std == 0 or noise_scale == 0 it is the identity on integer end points; otherwise the
// caller applies it per frame (rain_rendering_b200/streaks.py: apply_wind_noise, in NumPy like the reference, because
// the write-back makes frames depend on each other) -- src_index tells it which simulator record each output came from.
extern "C" int rr_host_assemble_batch(int n_frames, const rr_streak_rec *const *sim, const int32_t *n_sim, const uint32_t *seeds,
                                      int W, int H, const double *db_ratios, int n_ratios, double noise_std, double noise_scale,
                                      rr_streak_rec *out, int64_t out_cap, int32_t *offsets, int32_t *src_index) {
    if (n_frames < 0 || !n_sim || !seeds || !offsets || (n_frames > 0 && !sim) || (out_cap > 0 && !out) || (n_ratios > 0 && !db_ratios)) return RR_ERR_ARG;
    const int m = H > W ? H : W, nr = n_ratios < 4 ? n_ratios : 4;
    int64_t o = 0;
    offsets[0] = 0;
    MT mt;
    for (int f = 0; f < n_frames; f++) {
        const rr_streak_rec *S = sim[f];
        if (n_sim[f] < 0 || (n_sim[f] > 0 && !S)) return RR_ERR_ARG;
        mt.seed(seeds[f]);
        for (int i = 0; i < n_sim[f]; i++) {
            const rr_streak_rec &r = S[i];
            const bool ok_w = 1 <= r.max_width && r.max_width < m, ok_l = 1 <= r.length && r.length < m;
            const bool ins = 0 <= r.ip1m[0] && r.ip1m[0] < W && 0 <= r.ip1m[1] && r.ip1m[1] < H;
            const bool ine = 0 <= r.ip2m[0] && r.ip2m[0] < W && 0 <= r.ip2m[1] && r.ip2m[1] < H;
            if (!(ok_w && ok_l && (ins || ine))) continue;
            if (o >= out_cap) return RR_ERR_CAPACITY;
            rr_streak_rec &d = out[o];
            d = r;
            d.ip1[0] = r.ip1m[0]; d.ip1[1] = r.ip1m[1]; d.ip2[0] = r.ip2m[0]; d.ip2[1] = r.ip2m[1];
            int b = 0;
            while (b < nr && !(r.ratio < db_ratios[b])) b++;
            d.tex_idx = (uint8_t)(10 * b + (int)mt.bounded(9));
            d.noise_deg = r.type != 0 ? (0.0 + noise_std * mt.legacy_gauss()) * noise_scale : 0.0;
            if (src_index) src_index[o] = i;
            o++;
        }
        if (o > 0x7fffffff) return RR_ERR_CAPACITY;
        offsets[f + 1] = (int32_t)o;
    }
    return RR_OK;
}

// FovComputation.compute_fov_plane_points (common/bad_weather.py:596-704) on the host: the header code of the device path
// (rr_fov_polygon, csrc/rr_streak_geom.h), for callers that want the polygon itself.
extern "C" int rr_host_fov_polygon(const rr_streak_rec *rec, double radius, double fov_deg, int rows, int cols, double *xy,
                                   int32_t *n_vertices) {
    if (!rec || !xy || !n_vertices || rows <= 0 || cols <= 0) return RR_ERR_ARG;
    double px[RR_FOV_N + 4], py[RR_FOV_N + 4];
    const int n = rr_fov_polygon(*rec, radius, fov_deg, rows, cols, px, py);
    for (int i = 0; i < n; i++) { xy[2 * i] = px[i]; xy[2 * i + 1] = py[i]; }
    *n_vertices = n;
    return RR_OK;
}
