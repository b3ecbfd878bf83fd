// Native loader of the particle simulator's XML output (no CUDA), exported through the C ABI.
//
// Replaces DBManager.load_streaks_from_xml of the reference (common/bad_weather.py:148-248): the
// reference parses the file with xml.etree and builds one Python object per streak with a dozen
// small NumPy calls each (KITTI data_object files hold 101 frames x thousands of streaks and are
// re-parsed for every weather); this reads the file once, tokenises it in place and fills
// rr_streak_rec records directly.
//
// Semantics kept from the reference:
//   * the root's children are frames (<i id t d rs>), a frame's children are streaks (<r ...>);
//     tag names are not checked (:192,200), unknown attributes are ignored
//   * per streak (:201-238): positions and diameters / render_scale, image y flipped with the
//     image height, world z negated, max_width = int(max(iw1, iw2)), ratio from the un-rounded
//     positions, positions rounded half-even, length = ceil(|ip1 - ip2|) from the rounded ones,
//     type from max_width (:99-106)
//   * a streak enters the frame's dict only if max_width >= 1 and length >= 1 (:238); a later
//     streak with the same pid replaces the earlier one IN PLACE (dict.update keeps the slot)
//   * frames are keyed by their id the same way (:241)
// Numbers are converted with strtod / strtol: correctly rounded like Python's float().
#include <errno.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../include/rain_b200.h"

extern "C" void rr_set_error(const char *msg);

struct rr_xml_particles {
    std::vector<rr_xml_frame> frames;
    std::vector<rr_streak_rec> records;
};

namespace {

struct Attr { const char *name; size_t nlen; const char *val; size_t vlen; };

struct Parser {
    const char *p, *end;
    std::string err;
    bool fail(const char *what, const char *at) {
        char buf[160];
        snprintf(buf, sizeof(buf), "%s at byte %lld", what, (long long)(at - base));
        err = buf;
        return false;
    }
    const char *base;
    static bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
    static bool is_name(char c) { return !(is_space(c) || c == '=' || c == '>' || c == '/' || c == '<' || c == '"' || c == '\''); }
    void skip_space() { while (p < end && is_space(*p)) p++; }
    bool starts(const char *s) { size_t n = strlen(s); return (size_t)(end - p) >= n && memcmp(p, s, n) == 0; }
    bool skip_until(const char *s) {
        size_t n = strlen(s);
        for (; p + n <= end; p++)
            if (memcmp(p, s, n) == 0) { p += n; return true; }
        return false;
    }
    // Next tag.  kind: 0 = open, 1 = close, 2 = self-closing, -1 = end of input.  Text, comments,
    // processing instructions and declarations between tags are skipped.
    bool next_tag(int *kind, std::vector<Attr> *attrs) {
        for (;;) {
            while (p < end && *p != '<') p++;
            if (p >= end) { *kind = -1; return true; }
            const char *at = p;
            if (starts("<!--")) { if (!skip_until("-->")) return fail("unterminated comment", at); continue; }
            if (starts("<?")) { if (!skip_until("?>")) return fail("unterminated processing instruction", at); continue; }
            if (starts("<![CDATA[")) { if (!skip_until("]]>")) return fail("unterminated CDATA", at); continue; }
            if (starts("<!")) { if (!skip_until(">")) return fail("unterminated declaration", at); continue; }
            p++;
            if (p < end && *p == '/') {
                if (!skip_until(">")) return fail("unterminated end tag", at);
                *kind = 1;
                return true;
            }
            if (p >= end || !is_name(*p)) return fail("malformed tag", at);
            while (p < end && is_name(*p)) p++;
            attrs->clear();
            for (;;) {
                skip_space();
                if (p >= end) return fail("unterminated tag", at);
                if (*p == '>') { p++; *kind = 0; return true; }
                if (*p == '/') {
                    if (p + 1 < end && p[1] == '>') { p += 2; *kind = 2; return true; }
                    return fail("malformed tag", at);
                }
                Attr a;
                a.name = p;
                while (p < end && is_name(*p)) p++;
                a.nlen = (size_t)(p - a.name);
                if (a.nlen == 0) return fail("malformed attribute", at);
                skip_space();
                if (p >= end || *p != '=') return fail("attribute without value", at);
                p++;
                skip_space();
                if (p >= end || (*p != '"' && *p != '\'')) return fail("unquoted attribute value", at);
                char q = *p++;
                a.val = p;
                while (p < end && *p != q) p++;
                if (p >= end) return fail("unterminated attribute value", at);
                a.vlen = (size_t)(p - a.val);
                p++;
                attrs->push_back(a);
            }
        }
    }
};

// np.linalg.norm of a 2-vector = sqrt(ddot(x, x)); the BLAS kernel NumPy calls accumulates with fused
// multiply-adds on every x86-64 host since Haswell: sqrt(fma(y, y, x * x)).  Measured against NumPy 2.3 /
// OpenBLAS in the build container (20000 / 20000 identical); a host without FMA would give sqrt(x*x + y*y),
// at most one ulp away -- it only enters `ratio`, which picks the texture bucket (bad_weather.py:228-233,251).
inline double norm2_np(double x, double y) { return sqrt(fma(y, y, x * x)); }

const Attr *find_attr(const std::vector<Attr> &attrs, const char *name) {
    size_t n = strlen(name);
    const Attr *hit = nullptr;
    for (const Attr &a : attrs)
        if (a.nlen == n && memcmp(a.name, name, n) == 0) hit = &a;      // the last duplicate wins, like a dict
    return hit;
}

// float(str): leading/trailing blanks allowed, nothing else may follow the number
bool to_double(const char *s, size_t n, double *out) {
    char buf[96];
    if (n == 0 || n >= sizeof(buf)) return false;
    memcpy(buf, s, n);
    buf[n] = 0;
    char *e = nullptr;
    errno = 0;
    double v = strtod(buf, &e);
    if (e == buf) return false;
    while (*e == ' ' || *e == '\t' || *e == '\n' || *e == '\r') e++;
    if (*e) return false;
    *out = v;
    return true;
}

bool to_int(const char *s, size_t n, long long *out) {
    char buf[64];
    if (n == 0 || n >= sizeof(buf)) return false;
    memcpy(buf, s, n);
    buf[n] = 0;
    char *e = nullptr;
    errno = 0;
    long long v = strtoll(buf, &e, 10);
    if (e == buf || errno == ERANGE) return false;
    while (*e == ' ' || *e == '\t' || *e == '\n' || *e == '\r') e++;
    if (*e) return false;
    *out = v;
    return true;
}

// "[a;b;c]" -> k doubles: attr[1:-1].split(';') (bad_weather.py:202-207)
bool to_vec(const Attr *a, int k, double *out) {
    if (!a || a->vlen < 2) return false;
    const char *s = a->val + 1, *e = a->val + a->vlen - 1;
    for (int i = 0; i < k; i++) {
        const char *t = s;
        while (t < e && *t != ';') t++;
        if (!to_double(s, (size_t)(t - s), &out[i])) return false;
        if (i + 1 < k) { if (t >= e) return false; s = t + 1; }
        else if (t != e) return false;
    }
    return true;
}

bool attr_double(const std::vector<Attr> &attrs, const char *name, double *out) {
    const Attr *a = find_attr(attrs, name);
    return a && to_double(a->val, a->vlen, out);
}

bool attr_int(const std::vector<Attr> &attrs, const char *name, long long *out) {
    const Attr *a = find_attr(attrs, name);
    return a && to_int(a->val, a->vlen, out);
}

// one <r .../> -> record; keep = the reference's "max_width >= 1 and length >= 1"
bool make_record(const std::vector<Attr> &attrs, int render_scale, int H, rr_streak_rec *r, bool *keep) {
    long long pid;
    double wp1[3], wp2[3], ip1[2], ip2[2], iw1, iw2, wd;
    if (!attr_int(attrs, "pid", &pid)) return false;
    if (!to_vec(find_attr(attrs, "wp1"), 3, wp1) || !to_vec(find_attr(attrs, "wp2"), 3, wp2)) return false;
    if (!attr_double(attrs, "wd1", &wd) || !attr_double(attrs, "wd2", &wd)) return false;     // read (and required) by the reference
    if (!to_vec(find_attr(attrs, "ip1"), 2, ip1) || !to_vec(find_attr(attrs, "ip2"), 2, ip2)) return false;
    if (!attr_double(attrs, "iw1", &iw1) || !attr_double(attrs, "iw2", &iw2)) return false;
    const double rs = (double)render_scale;
    for (int i = 0; i < 2; i++) { ip1[i] = ip1[i] / rs; ip2[i] = ip2[i] / rs; }      // :208-209
    iw1 = iw1 / rs; iw2 = iw2 / rs;                                                   // :210-211
    ip1[1] = H - ip1[1]; ip2[1] = H - ip2[1];                                         // :221-222
    wp1[2] *= -1; wp2[2] *= -1;                                                       // :223-224
    const double dx = fabs(ip1[0] - ip2[0]), dy = fabs(ip1[1] - ip2[1]);
    const double mw_f = iw1 >= iw2 ? iw1 : iw2;                                       // max(iw1, iw2)
    if (!(mw_f == mw_f) || fabs(mw_f) > 9e18) return false;                           // int(nan) / int(inf) raise
    const long long max_width = (long long)mw_f;                                      // :226 truncation
    const double nrm = norm2_np(dx, dy);
    const double cos_theta = 0 * (dx / nrm) + -1 * (-(dy / nrm));                     // :228-231
    const double ratio = (double)max_width / (dy / cos_theta);                        // :232-233
    for (int i = 0; i < 2; i++)                                                       // a coordinate no image has: refuse it rather than
        if (!(fabs(ip1[i]) < 1e9) || !(fabs(ip2[i]) < 1e9)) return false;             // overflow the integer conversions below (NaN included)
    const long long x1 = (long long)nearbyint(ip1[0]), y1 = (long long)nearbyint(ip1[1]);   // :234-235 half-even
    const long long x2 = (long long)nearbyint(ip2[0]), y2 = (long long)nearbyint(ip2[1]);
    const double ex = (double)(x1 - x2), ey = (double)(y1 - y2);
    const long long length = (long long)ceil(sqrt(ex * ex + ey * ey));                // :236
    memset(r, 0, sizeof(*r));
    for (int i = 0; i < 3; i++) { r->wp1[i] = wp1[i]; r->wp2[i] = wp2[i]; }
    r->iw1 = iw1; r->iw2 = iw2; r->noise_deg = 0; r->ratio = ratio;
    r->ip1[0] = r->ip1m[0] = (int32_t)x1; r->ip1[1] = r->ip1m[1] = (int32_t)y1;
    r->ip2[0] = r->ip2m[0] = (int32_t)x2; r->ip2[1] = r->ip2m[1] = (int32_t)y2;
    r->max_width = (int32_t)max_width; r->length = (int32_t)length; r->pid = (int32_t)pid;
    r->type = max_width >= 4 ? 0 : (max_width > 1 ? 1 : 2);                           // :99-106
    *keep = max_width >= 1 && length >= 1;                                            // :238
    return true;
}

bool read_file(const char *path, std::vector<char> *buf, std::string *err) {
    FILE *f = fopen(path, "rb");
    if (!f) { *err = std::string("cannot open ") + path + ": " + strerror(errno); return false; }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (n < 0) { fclose(f); *err = "cannot size the file"; return false; }
    buf->resize((size_t)n);
    size_t got = n ? fread(buf->data(), 1, (size_t)n, f) : 0;
    fclose(f);
    if (got != (size_t)n) { *err = "short read"; return false; }
    return true;
}

}  // namespace

extern "C" int rr_host_load_particles_xml(const char *path, int render_scale, int W, int H, rr_xml_particles **out) {
    (void)W;
    if (!path || !out || render_scale <= 0) { rr_set_error("rr_host_load_particles_xml: bad arguments"); return RR_ERR_ARG; }
    std::vector<char> buf;
    std::string err;
    if (!read_file(path, &buf, &err)) { rr_set_error(("rr_host_load_particles_xml: " + err).c_str()); return RR_ERR_ARG; }
    Parser ps;
    ps.base = ps.p = buf.data();
    ps.end = buf.data() + buf.size();
    std::vector<Attr> attrs;
    // frames in dict order (first insertion of an id keeps the slot, a later one replaces the content)
    struct FrameTmp { rr_xml_frame hdr; std::vector<rr_streak_rec> recs; };
    std::vector<FrameTmp> frames;
    std::unordered_map<long long, size_t> frame_slot;
    int depth = 0;
    bool have_root = false, root_closed = false;
    FrameTmp cur;
    std::unordered_map<long long, size_t> pid_slot;
    bool in_frame = false;
    for (;;) {
        int kind;
        const char *at = ps.p;
        if (!ps.next_tag(&kind, &attrs)) {
            rr_set_error(("rr_host_load_particles_xml: " + std::string(path) + ": " + ps.err +
                          " (corrupted particles simulation file? delete it and re-run the simulation)").c_str());
            return RR_ERR_ARG;
        }
        if (kind < 0) break;
        if (kind == 1) {               // end tag
            if (depth == 0) { ps.fail("unbalanced end tag", at); goto malformed; }
            if (depth == 2 && in_frame) {
                auto it = frame_slot.find(cur.hdr.id);
                if (it == frame_slot.end()) { frame_slot[cur.hdr.id] = frames.size(); frames.push_back(std::move(cur)); }
                else frames[it->second] = std::move(cur);
                in_frame = false;
            }
            depth--;
            if (depth == 0) root_closed = true;
            continue;
        }
        // open or self-closing element at depth + 1
        const int level = depth + 1;
        if (level == 1) {
            if (have_root) { ps.fail("junk after the document element", at); goto malformed; }
            have_root = true;
        } else if (level == 2) {
            long long id, t, d, rs;
            if (!attr_int(attrs, "id", &id) || !attr_int(attrs, "t", &t) || !attr_int(attrs, "d", &d) || !attr_int(attrs, "rs", &rs)) {
                ps.fail("frame element without integer id / t / d / rs attributes", at);
                goto malformed;
            }
            cur = FrameTmp();
            cur.hdr.id = (int32_t)id; cur.hdr.exposure_t = (int32_t)t; cur.hdr.start_d = (int32_t)d; cur.hdr.streaks_count = (int32_t)rs;
            pid_slot.clear();
            in_frame = true;
            if (kind == 2) {           // <i ... /> : a frame without streaks
                auto it = frame_slot.find(cur.hdr.id);
                if (it == frame_slot.end()) { frame_slot[cur.hdr.id] = frames.size(); frames.push_back(std::move(cur)); }
                else frames[it->second] = std::move(cur);
                in_frame = false;
            }
        } else if (level == 3 && in_frame) {
            rr_streak_rec r;
            bool keep = false;
            if (!make_record(attrs, render_scale, H, &r, &keep)) { ps.fail("streak element with missing or malformed attributes", at); goto malformed; }
            if (keep) {
                auto it = pid_slot.find(r.pid);
                if (it == pid_slot.end()) { pid_slot[r.pid] = cur.recs.size(); cur.recs.push_back(r); }
                else cur.recs[it->second] = r;
            }
        }
        if (kind == 0) depth++;
    }
    if (!have_root || depth != 0 || !root_closed) {
        if (have_root && depth == 0 && !root_closed) {
            // a self-closing root (<camera/>): no frames
        } else {
            ps.fail("unexpected end of file", ps.end);
            goto malformed;
        }
    }
    {
        rr_xml_particles *res = new rr_xml_particles();
        size_t total = 0;
        for (auto &f : frames) total += f.recs.size();
        res->records.reserve(total);
        for (auto &f : frames) {
            f.hdr.first = (int64_t)res->records.size();
            f.hdr.count = (int64_t)f.recs.size();
            res->records.insert(res->records.end(), f.recs.begin(), f.recs.end());
            res->frames.push_back(f.hdr);
        }
        *out = res;
        return RR_OK;
    }
malformed:
    rr_set_error(("rr_host_load_particles_xml: " + std::string(path) + ": " + ps.err +
                  " (corrupted particles simulation file? delete it and re-run the simulation)").c_str());
    return RR_ERR_ARG;
}

extern "C" int rr_host_particles_info(const rr_xml_particles *p, int32_t *n_frames, int64_t *n_records) {
    if (!p) { rr_set_error("rr_host_particles_info: NULL handle"); return RR_ERR_ARG; }
    if (n_frames) *n_frames = (int32_t)p->frames.size();
    if (n_records) *n_records = (int64_t)p->records.size();
    return RR_OK;
}

extern "C" int rr_host_particles_copy(const rr_xml_particles *p, rr_xml_frame *frames, rr_streak_rec *records) {
    if (!p) { rr_set_error("rr_host_particles_copy: NULL handle"); return RR_ERR_ARG; }
    if (frames && !p->frames.empty()) memcpy(frames, p->frames.data(), p->frames.size() * sizeof(rr_xml_frame));
    if (records && !p->records.empty()) memcpy(records, p->records.data(), p->records.size() * sizeof(rr_streak_rec));
    return RR_OK;
}

extern "C" void rr_host_free_particles(rr_xml_particles *p) { delete p; }

extern "C" void rr_host_norm2(int n, const double *x, const double *y, double *out) {
    for (int i = 0; i < n; i++) out[i] = norm2_np(x[i], y[i]);
}
