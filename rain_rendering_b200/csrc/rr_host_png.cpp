// Native PNG codec of the frame pipeline (no CUDA): the reference decodes every frame with cv2.imread
// (common/generator.py:352,360) and encodes two PNGs per frame through matplotlib (generator.py:466-467);
// once rendering takes 0.1 ms per frame those ~115 ms of single-threaded libpng work per frame are the whole
// run time.  This codec does the same job on a pool of native threads, straight from / into the page-locked
// batch buffers of rr_submit_frames:
//   * decode: non-interlaced 8-bit gray / gray+alpha / RGB / RGBA and 16-bit gray / RGB, exactly the arrays
//     cv2.imread(path) (3-channel BGR uint8, alpha dropped, 16-bit reduced to its high byte) and
//     cv2.imread(path, IMREAD_UNCHANGED).astype(float32) / 256 (the depth path, generator.py:360-365) give;
//     anything else (palette, interlace, sub-byte depths) reports RR_PNG_UNSUPPORTED and the caller falls back
//   * encode: 8-bit RGB from the BGR batch buffer, and the rain mask min/max-normalised to 16-bit gray (the
//     drop-in's stand-in for plt.imsave of a 2-D array); one fixed filter per image (Sub) and, at the default
//     level 1, Huffman-only deflate: about a third of cv2.imwrite's time per file.
// zlib (inflate / deflate / crc32) is the only dependency.
#include <errno.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#include <atomic>
#include <string>
#include <thread>
#include <vector>
#include "../../include/rain_b200.h"
#include "rr_host_deflate.h"
#include "rr_host_inflate.h"
#include "rr_viridis.h"

extern "C" void rr_set_error(const char *msg);

namespace {

const unsigned char kSig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};

uint32_t be32(const unsigned char *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
void put32(unsigned char *p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }

bool read_file(const char *path, std::vector<unsigned char> *buf) {
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (n < 0) { fclose(f); return false; }
    buf->resize((size_t)n);
    size_t got = n ? fread(buf->data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return got == (size_t)n;
}

struct Image {
    int w = 0, h = 0, channels = 0, depth = 0;     // depth: bits per sample (8 or 16)
    std::vector<unsigned char> px;                  // unfiltered scanlines, each (stride + 1) bytes: filter byte, then w * channels * depth/8 sample bytes (big-endian)
    size_t stride = 0;
    const unsigned char *row(int y) const { return px.data() + (stride + 1) * (size_t)y + 1; }
};

int paeth(int a, int b, int c) {
    int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// Where decode() delivers the pixels: row by row, right after a scanline is unfiltered (it is still in L1 then), into the
// caller's batch buffer.  SINK_NONE keeps them in Image::px only.
enum { SINK_NONE = 0, SINK_BGR8 = 1, SINK_DEPTH_U16 = 2, SINK_DEPTH_F32 = 3 };
struct RowSink { int kind = SINK_NONE; void *dst = nullptr; };

// cv2.imread(path) semantics for one scanline: BGR uint8, gray replicated, alpha dropped, 16-bit -> high byte
inline void row_to_bgr8(int w, int ch, int step, const unsigned char *s, uint8_t *d) {
    if (ch == 1 || ch == 2) {
        for (int i = 0; i < w; i++) { uint8_t g = s[(size_t)i * ch * step]; d[3 * i] = d[3 * i + 1] = d[3 * i + 2] = g; }
    } else {
        for (int i = 0; i < w; i++) {
            const unsigned char *p = s + (size_t)i * ch * step;
            d[3 * i] = p[2 * step]; d[3 * i + 1] = p[step]; d[3 * i + 2] = p[0];
        }
    }
}
// cv2.imread(path, IMREAD_UNCHANGED) of a single-channel scanline: uint16 samples, or .astype(np.float32) / 256.
inline void row_to_u16(int w, int depth, const unsigned char *s, uint16_t *d) {
    if (depth == 16) for (int i = 0; i < w; i++) d[i] = (uint16_t)((s[2 * i] << 8) | s[2 * i + 1]);
    else for (int i = 0; i < w; i++) d[i] = s[i];
}
inline void row_to_f32(int w, int depth, const unsigned char *s, float *d) {
    if (depth == 16) for (int i = 0; i < w; i++) d[i] = (float)((s[2 * i] << 8) | s[2 * i + 1]) / 256.f;
    else for (int i = 0; i < w; i++) d[i] = (float)s[i] / 256.f;
}

// -> RR_OK, RR_PNG_UNSUPPORTED (valid PNG this codec does not handle) or RR_ERR_ARG (not a readable PNG)
// want_w / want_h > 0: the size the caller's buffer was made for -- anything else returns RR_PNG_SIZE right after the
// header, before a single byte is allocated for the pixels (a 60-byte file may claim 65536 x 65536 RGBA16 = 34 GB)
int decode(const char *path, Image *img, bool header_only, int want_w = 0, int want_h = 0, const RowSink *sink = nullptr) {
    std::vector<unsigned char> file;
    if (!read_file(path, &file)) return RR_ERR_ARG;
    if (file.size() < 8 + 25 || memcmp(file.data(), kSig, 8) != 0) return RR_ERR_ARG;
    size_t pos = 8;
    bool have_hdr = false;
    int color = 0, interlace = 0;
    std::vector<unsigned char> raw, zdata;
    size_t stride = 0;
    int ret = RR_ERR_ARG;
    bool have_end = false;
    while (pos + 12 <= file.size()) {
        uint32_t len = be32(&file[pos]);
        const unsigned char *type = &file[pos + 4], *data = &file[pos + 8];
        if (pos + 12 + (size_t)len > file.size()) break;
        if (!memcmp(type, "IHDR", 4)) {
            if (len != 13 || have_hdr) break;
            img->w = (int)be32(data); img->h = (int)be32(data + 4);
            img->depth = data[8]; color = data[9]; interlace = data[12];
            if (img->w <= 0 || img->h <= 0 || img->w > 65536 || img->h > 65536) break;
            have_hdr = true;
            img->channels = color == 0 ? 1 : color == 2 ? 3 : color == 4 ? 2 : color == 6 ? 4 : 0;
            if (header_only) { if (color == 3) img->channels = 3; return RR_OK; }
            if (want_w > 0 && (img->w != want_w || img->h != want_h)) { ret = RR_PNG_SIZE; break; }
            if ((uint64_t)img->w * (uint64_t)img->h > ((uint64_t)1 << 28)) { ret = RR_PNG_UNSUPPORTED; break; }     // 268 Mpx: not a camera frame
            if (img->channels == 0 || interlace != 0 || (img->depth != 8 && img->depth != 16) || data[10] != 0 || data[11] != 0) {
                ret = RR_PNG_UNSUPPORTED;
                break;
            }
            stride = (size_t)img->w * img->channels * (img->depth / 8);
            if ((stride + 1) * (size_t)img->h + 1 > 0xfffffff0u) { ret = RR_PNG_UNSUPPORTED; break; }              // zlib counts in 32 bits
            zdata.reserve(file.size());
        } else if (!memcmp(type, "IDAT", 4)) {
            if (!have_hdr) break;
            zdata.insert(zdata.end(), data, data + len);          // the stream may be cut into several chunks
        } else if (!memcmp(type, "IEND", 4)) {
            have_end = true;
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (ret == RR_ERR_ARG && have_hdr && have_end && !zdata.empty()) {
        // the whole zlib stream and its exact decoded size are known: one pass of the library's own decoder
        // (rr_host_inflate.h); whatever it refuses gets a second opinion from zlib, so a valid file is never lost
        const size_t zlen = zdata.size(), need = (stride + 1) * (size_t)img->h;
        zdata.resize(zlen + 16, 0);                                // readable slack for the 8-byte bit-buffer refills
        raw.resize(need + 1);
        if (rr_inflate::zlib_decompress(zdata.data(), zlen, raw.data(), need)) ret = RR_OK;
        else {
            z_stream zs;
            memset(&zs, 0, sizeof(zs));
            if (inflateInit(&zs) == Z_OK) {
                zs.next_in = zdata.data(); zs.avail_in = (uInt)zlen;
                zs.next_out = raw.data(); zs.avail_out = (uInt)raw.size();   // one spare byte: the stream must END (Adler-32 verified), not just fill the image
                if (inflate(&zs, Z_FINISH) == Z_STREAM_END && zs.total_out == need) ret = RR_OK;
                inflateEnd(&zs);
            }
        }
    }
    if (ret != RR_OK) return ret;
    const int sink_kind = sink ? sink->kind : SINK_NONE;
    if ((sink_kind == SINK_DEPTH_U16 || sink_kind == SINK_DEPTH_F32) && img->channels != 1) return RR_PNG_UNSUPPORTED;
    // unfilter IN PLACE (each scanline keeps its filter byte in front): no second buffer, and the Sub filter -- what
    // OpenCV's writer and this library's own use throughout -- runs as bpp independent running sums
    const int bpp = img->channels * (img->depth / 8);
    img->stride = stride;
    unsigned char *prev = nullptr;
    for (int y = 0; y < img->h; y++) {
        unsigned char *row = raw.data() + (stride + 1) * (size_t)y + 1;
        const int ft = row[-1];
        // Sub / None scanlines of the two layouts the frame pipeline reads (8-bit RGB, 16-bit gray) go to the sink in ONE pass:
        // the running sums are written straight into the batch buffer in its sample order, and the scanline stays filtered
        // in `raw` -- allowed when the next scanline does not look at this one (Up, Average and Paeth do)
        const bool next_reads_this = y + 1 < img->h && raw[(stride + 1) * (size_t)(y + 1)] >= 2;
        if (ft <= 1 && !next_reads_this && sink_kind == SINK_BGR8 && img->channels == 3 && img->depth == 8) {
            uint8_t *d = (uint8_t *)sink->dst + (size_t)y * img->w * 3;
            if (ft == 1) {
                unsigned a = 0, b = 0, c = 0;
                for (size_t i = 0; i + 3 <= stride; i += 3) { a += row[i]; b += row[i + 1]; c += row[i + 2]; d[i] = (uint8_t)c; d[i + 1] = (uint8_t)b; d[i + 2] = (uint8_t)a; }
            } else {
                for (size_t i = 0; i + 3 <= stride; i += 3) { d[i] = row[i + 2]; d[i + 1] = row[i + 1]; d[i + 2] = row[i]; }
            }
            prev = row;
            continue;
        }
        if (ft <= 1 && !next_reads_this && sink_kind == SINK_DEPTH_U16 && img->channels == 1 && img->depth == 16) {
            uint16_t *d = (uint16_t *)sink->dst + (size_t)y * img->w;
            if (ft == 1) {
                unsigned a = 0, b = 0;
                for (int i = 0; i < img->w; i++) { a += row[2 * i]; b += row[2 * i + 1]; d[i] = (uint16_t)(((a & 255u) << 8) | (b & 255u)); }
            } else row_to_u16(img->w, 16, row, d);
            prev = row;
            continue;
        }
        switch (ft) {
            case 0: break;
            case 1:
                if (bpp == 3) {
                    unsigned a = 0, b = 0, c = 0;
                    size_t i = 0;
                    for (; i + 3 <= stride; i += 3) { a += row[i]; b += row[i + 1]; c += row[i + 2]; row[i] = (unsigned char)a; row[i + 1] = (unsigned char)b; row[i + 2] = (unsigned char)c; }
                } else if (bpp == 2) {
                    unsigned a = 0, b = 0;
                    for (size_t i = 0; i + 2 <= stride; i += 2) { a += row[i]; b += row[i + 1]; row[i] = (unsigned char)a; row[i + 1] = (unsigned char)b; }
                } else if (bpp == 4) {
                    unsigned a = 0, b = 0, c = 0, d = 0;
                    for (size_t i = 0; i + 4 <= stride; i += 4) {
                        a += row[i]; b += row[i + 1]; c += row[i + 2]; d += row[i + 3];
                        row[i] = (unsigned char)a; row[i + 1] = (unsigned char)b; row[i + 2] = (unsigned char)c; row[i + 3] = (unsigned char)d;
                    }
                } else {
                    for (size_t i = bpp; i < stride; i++) row[i] = (unsigned char)(row[i] + row[i - bpp]);
                }
                break;
            case 2:
                if (prev) for (size_t i = 0; i < stride; i++) row[i] = (unsigned char)(row[i] + prev[i]);
                break;
            case 3:
                for (size_t i = 0; i < stride; i++) {
                    int a = i >= (size_t)bpp ? row[i - bpp] : 0, b = prev ? prev[i] : 0;
                    row[i] = (unsigned char)(row[i] + ((a + b) >> 1));
                }
                break;
            case 4:
                for (size_t i = 0; i < stride; i++) {
                    int a = i >= (size_t)bpp ? row[i - bpp] : 0, b = prev ? prev[i] : 0, c = (prev && i >= (size_t)bpp) ? prev[i - bpp] : 0;
                    row[i] = (unsigned char)(row[i] + paeth(a, b, c));
                }
                break;
            default: return RR_ERR_ARG;
        }
        if (sink_kind == SINK_BGR8) row_to_bgr8(img->w, img->channels, img->depth / 8, row, (uint8_t *)sink->dst + (size_t)y * img->w * 3);
        else if (sink_kind == SINK_DEPTH_U16) row_to_u16(img->w, img->depth, row, (uint16_t *)sink->dst + (size_t)y * img->w);
        else if (sink_kind == SINK_DEPTH_F32) row_to_f32(img->w, img->depth, row, (float *)sink->dst + (size_t)y * img->w);
        prev = row;
    }
    img->px.swap(raw);                  // scanlines of stride + 1 bytes (filter byte first), unfiltered
    return RR_OK;
}

void chunk(std::vector<unsigned char> *out, const char *type, const unsigned char *data, size_t len) {
    unsigned char hdr[8];
    put32(hdr, (uint32_t)len);
    memcpy(hdr + 4, type, 4);
    out->insert(out->end(), hdr, hdr + 8);
    if (len) out->insert(out->end(), data, data + len);
    uLong c = crc32(0L, hdr + 4, 4);
    if (len) c = crc32(c, data, (uInt)len);
    unsigned char tail[4];
    put32(tail, (uint32_t)c);
    out->insert(out->end(), tail, tail + 4);
}

// filt: h filtered scanlines ((stride + 1) bytes each, filter-type byte first) -> PNG file.
// level 0 stores, 1 = this library's run + Huffman encoder (rr_host_deflate.h), 2..9 = zlib at that level.
int write_png(const char *path, const std::vector<unsigned char> &filt, int w, int h, int color, int depth, int level) {
    std::vector<unsigned char> comp;
    if (level == 1) rr_deflate::zlib_compress_fast(filt.data(), filt.size(), comp);
    else {
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (deflateInit2(&zs, level, Z_DEFLATED, 15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return RR_ERR_ARG;
        comp.resize(deflateBound(&zs, (uLong)filt.size()));
        zs.next_in = const_cast<unsigned char *>(filt.data()); zs.avail_in = (uInt)filt.size();
        zs.next_out = comp.data(); zs.avail_out = (uInt)comp.size();
        int r = deflate(&zs, Z_FINISH);
        comp.resize(zs.total_out);
        deflateEnd(&zs);
        if (r != Z_STREAM_END) return RR_ERR_ARG;
    }
    std::vector<unsigned char> out;
    out.reserve(comp.size() + 128);
    out.insert(out.end(), kSig, kSig + 8);
    unsigned char ihdr[13];
    put32(ihdr, (uint32_t)w); put32(ihdr + 4, (uint32_t)h);
    ihdr[8] = (unsigned char)depth; ihdr[9] = (unsigned char)color; ihdr[10] = ihdr[11] = ihdr[12] = 0;
    chunk(&out, "IHDR", ihdr, 13);
    chunk(&out, "IDAT", comp.data(), comp.size());
    chunk(&out, "IEND", nullptr, 0);
    std::string tmp = std::string(path) + ".part";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return RR_ERR_ARG;
    bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    ok = (fclose(f) == 0) && ok;
    if (!ok || rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); return RR_ERR_ARG; }
    return RR_OK;
}

// rows: h scanlines of `stride` bytes (samples already big-endian / RGB order); bpp = bytes per pixel
int encode(const char *path, const unsigned char *rows, int w, int h, int color, int depth, int bpp, int level) {
    const size_t stride = (size_t)w * bpp;
    std::vector<unsigned char> filt((stride + 1) * (size_t)h);
    const int ft = level > 0 ? 1 : 0;           // Sub: cheap, and what makes smooth images compressible; None when storing
    for (int y = 0; y < h; y++) {
        const unsigned char *s = rows + stride * (size_t)y;
        unsigned char *d = filt.data() + (stride + 1) * (size_t)y;
        d[0] = (unsigned char)ft;
        d++;
        if (ft == 0) memcpy(d, s, stride);
        else {
            for (int i = 0; i < bpp && (size_t)i < stride; i++) d[i] = s[i];
            for (size_t i = bpp; i < stride; i++) d[i] = (unsigned char)(s[i] - s[i - bpp]);
        }
    }
    return write_png(path, filt, w, h, color, depth, level);
}

// The reference's file formats (generator.py:466-467, plt.imsave): 8-bit RGBA, alpha 255.  Colour conversion, alpha and the
// Sub filter in one pass over the batch buffer.
//   image: from the BGR uint8 frame;  mask: from the colormap index through matplotlib's viridis table (rr_viridis.h)
int write_rgba(const char *path, const uint8_t *bgr, const uint8_t *idx8, int w, int h, int level) {
    const size_t stride = (size_t)w * 4;
    std::vector<unsigned char> filt((stride + 1) * (size_t)h);
    const bool sub = level > 0;
    for (int y = 0; y < h; y++) {
        unsigned char *d = filt.data() + (stride + 1) * (size_t)y;
        *d++ = sub ? 1 : 0;
        unsigned pr = 0, pg = 0, pb = 0, pa = 0;
        for (int x = 0; x < w; x++) {
            unsigned r, g, b;
            if (bgr) { const uint8_t *p = bgr + ((size_t)y * w + x) * 3; r = p[2]; g = p[1]; b = p[0]; }
            else { const uint8_t *c = rr_viridis_rgb[idx8[(size_t)y * w + x]]; r = c[0]; g = c[1]; b = c[2]; }
            d[4 * x] = (unsigned char)(r - pr); d[4 * x + 1] = (unsigned char)(g - pg); d[4 * x + 2] = (unsigned char)(b - pb);
            d[4 * x + 3] = (unsigned char)(255u - pa);
            if (sub) { pr = r; pg = g; pb = b; pa = 255u; }
        }
    }
    return write_png(path, filt, w, h, 6, 8, level);
}

int write_gray16(const char *path, const uint16_t *v, int w, int h, int level) {
    const size_t n = (size_t)w * h;
    std::vector<unsigned char> g(n * 2);
    for (size_t i = 0; i < n; i++) { g[2 * i] = (unsigned char)(v[i] >> 8); g[2 * i + 1] = (unsigned char)v[i]; }
    return encode(path, g.data(), w, h, 0, 16, 2, level);
}

int write_bgr8(const char *path, const uint8_t *bgr, int w, int h, int level) {
    std::vector<unsigned char> rgb((size_t)w * h * 3);
    for (size_t i = 0, n = (size_t)w * h; i < n; i++) { rgb[3 * i] = bgr[3 * i + 2]; rgb[3 * i + 1] = bgr[3 * i + 1]; rgb[3 * i + 2] = bgr[3 * i]; }
    return encode(path, rgb.data(), w, h, 2, 8, 3, level);
}

// the drop-in's mask file: (mask - min) / (max - min) * 65535 + 0.5 as 16-bit gray (zeros when the mask is flat)
int write_mask16(const char *path, const float *mask, int w, int h, int level) {
    const size_t n = (size_t)w * h;
    float lo = mask[0], hi = mask[0];
    for (size_t i = 1; i < n; i++) { lo = mask[i] < lo ? mask[i] : lo; hi = mask[i] > hi ? mask[i] : hi; }
    std::vector<unsigned char> g(n * 2);
    const double dlo = lo, range = (double)hi - (double)lo;
    for (size_t i = 0; i < n; i++) {
        double v = range > 0 ? ((double)mask[i] - dlo) / range : 0.0;
        unsigned q = (unsigned)(v * 65535.0 + 0.5);
        g[2 * i] = (unsigned char)(q >> 8); g[2 * i + 1] = (unsigned char)q;
    }
    return encode(path, g.data(), w, h, 0, 16, 2, level);
}

template <class F>
void parallel_for(int n, int n_threads, F fn) {
    if (n_threads > n) n_threads = n;
    if (n_threads <= 1) { for (int i = 0; i < n; i++) fn(i); return; }
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++)
        th.emplace_back([&]() { for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i); });
    for (auto &t : th) t.join();
}

}  // namespace

extern "C" int rr_host_png_info(const char *path, int32_t *w, int32_t *h, int32_t *channels, int32_t *bit_depth) {
    if (!path) { rr_set_error("rr_host_png_info: path is NULL"); return RR_ERR_ARG; }
    Image im;
    int r = RR_ERR_ARG;
    try { r = decode(path, &im, true); } catch (...) { r = RR_ERR_ARG; }
    if (r != RR_OK) { rr_set_error((std::string("rr_host_png_info: cannot read ") + path).c_str()); return r; }
    if (w) *w = im.w;
    if (h) *h = im.h;
    if (channels) *channels = im.channels;
    if (bit_depth) *bit_depth = im.depth;
    return RR_OK;
}

// Decodes n frames into the batch buffers (frame i at bgr + i * 3 * Wi * Hi and depth + i * Wd * Hd); a NULL path
// list skips that kind.  status[i] receives RR_OK, RR_PNG_UNSUPPORTED / RR_ERR_ARG (file i must go through the
// caller's fallback decoder) or RR_PNG_SIZE (decodable, but not the expected size).  Returns the number of
// frames that are not RR_OK.  depth_u16 != 0: depth points to uint16 samples (rr_frame_io RR_DEPTH_U16_256).
static int read_batch(int n, const char *const *image_paths, const char *const *depth_paths, uint8_t *bgr, int Wi, int Hi,
                      void *depth, int depth_u16, int Wd, int Hd, int n_threads, int32_t *status) {
    if (n < 0 || !status || (image_paths && !bgr) || (depth_paths && !depth)) { rr_set_error("rr_host_png_read_batch: bad arguments"); return RR_ERR_ARG; }
    for (int i = 0; i < n; i++) status[i] = RR_OK;
    const int jobs = 2 * n;
    parallel_for(jobs, n_threads, [&](int j) {
        const int i = j >> 1, kind = j & 1;
        const char *const *paths = kind ? depth_paths : image_paths;
        if (!paths || !paths[i]) return;
        int r = RR_ERR_ARG;
        try {
            Image im;
            RowSink sink;
            if (kind == 0) { sink.kind = SINK_BGR8; sink.dst = bgr + (size_t)i * 3 * Wi * Hi; }
            else if (depth_u16) { sink.kind = SINK_DEPTH_U16; sink.dst = (uint16_t *)depth + (size_t)i * Wd * Hd; }
            else { sink.kind = SINK_DEPTH_F32; sink.dst = (float *)depth + (size_t)i * Wd * Hd; }
            r = decode(paths[i], &im, false, kind ? Wd : Wi, kind ? Hd : Hi, &sink);
        } catch (...) { r = RR_ERR_ARG; }              // std::bad_alloc and friends must not cross a thread / the C ABI
        if (r != RR_OK) __atomic_store_n(&status[i], (int32_t)r, __ATOMIC_RELAXED);     // two writers at most (image, depth), both storing a failure code
    });
    int bad = 0;
    for (int i = 0; i < n; i++) bad += status[i] != RR_OK;
    return bad;
}

extern "C" int rr_host_png_read_batch(int n, const char *const *image_paths, const char *const *depth_paths, uint8_t *bgr, int Wi, int Hi,
                                      float *depth, int Wd, int Hd, int n_threads, int32_t *status) {
    return read_batch(n, image_paths, depth_paths, bgr, Wi, Hi, depth, 0, Wd, Hd, n_threads, status);
}

extern "C" int rr_host_png_read_batch_u16(int n, const char *const *image_paths, const char *const *depth_paths, uint8_t *bgr, int Wi, int Hi,
                                          uint16_t *depth, int Wd, int Hd, int n_threads, int32_t *status) {
    return read_batch(n, image_paths, depth_paths, bgr, Wi, Hi, depth, 1, Wd, Hd, n_threads, status);
}

// The reference's output files (generator.py:466-467): both 8-bit RGBA like plt.imsave writes them -- the rainy image from
// the uint8 BGR frame, the rain mask from its colormap index (rr_frame_io.out_mask_idx8) through matplotlib's viridis
// table.  Either list may be NULL.  Returns the number of files that failed.
extern "C" int rr_host_png_write_batch_rgba(int n, const char *const *image_paths, const uint8_t *bgr, const char *const *mask_paths,
                                            const uint8_t *mask_idx8, int W, int H, int level, int n_threads) {
    if (n < 0 || W <= 0 || H <= 0 || (image_paths && !bgr) || (mask_paths && !mask_idx8) || level < 0 || level > 9) {
        rr_set_error("rr_host_png_write_batch_rgba: bad arguments");
        return RR_ERR_ARG;
    }
    std::atomic<int> bad(0);
    parallel_for(2 * n, n_threads, [&](int j) {
        const int i = j >> 1, kind = j & 1;
        int r = RR_OK;
        try {
            if (kind == 0 && image_paths && image_paths[i]) r = write_rgba(image_paths[i], bgr + (size_t)i * 3 * W * H, nullptr, W, H, level);
            if (kind == 1 && mask_paths && mask_paths[i]) r = write_rgba(mask_paths[i], nullptr, mask_idx8 + (size_t)i * W * H, W, H, level);
        } catch (...) { r = RR_ERR_ARG; }
        if (r != RR_OK) bad.fetch_add(1);
    });
    if (bad.load()) rr_set_error("rr_host_png_write_batch_rgba: some files could not be written");
    return bad.load();
}

// Compact output files: 8-bit RGB image, 16-bit gray mask from the normalised uint16 mask (rr_frame_io.out_mask_u16).
extern "C" int rr_host_png_write_batch_u16(int n, const char *const *image_paths, const uint8_t *bgr, const char *const *mask_paths,
                                           const uint16_t *mask_u16, int W, int H, int level, int n_threads) {
    if (n < 0 || W <= 0 || H <= 0 || (image_paths && !bgr) || (mask_paths && !mask_u16) || level < 0 || level > 9) {
        rr_set_error("rr_host_png_write_batch_u16: bad arguments");
        return RR_ERR_ARG;
    }
    std::atomic<int> bad(0);
    parallel_for(2 * n, n_threads, [&](int j) {
        const int i = j >> 1, kind = j & 1;
        int r = RR_OK;
        try {
            if (kind == 0 && image_paths && image_paths[i]) r = write_bgr8(image_paths[i], bgr + (size_t)i * 3 * W * H, W, H, level);
            if (kind == 1 && mask_paths && mask_paths[i]) r = write_gray16(mask_paths[i], mask_u16 + (size_t)i * W * H, W, H, level);
        } catch (...) { r = RR_ERR_ARG; }
        if (r != RR_OK) bad.fetch_add(1);
    });
    if (bad.load()) rr_set_error("rr_host_png_write_batch_u16: some files could not be written");
    return bad.load();
}

// Frames finished zlib streams (made on the GPU, rr_frame_io.out_png_*) as 8-bit RGBA PNG files.
extern "C" int rr_host_png_write_streams(int n, const char *const *paths, const uint8_t *streams, size_t stride, const uint32_t *sizes,
                                         int W, int H, int n_threads) {
    if (n < 0 || W <= 0 || H <= 0 || (n > 0 && (!paths || !streams || !sizes))) { rr_set_error("rr_host_png_write_streams: bad arguments"); return RR_ERR_ARG; }
    std::atomic<int> bad(0);
    parallel_for(n, n_threads, [&](int i) {
        int r = RR_ERR_ARG;
        try {
            if (paths[i] && sizes[i] >= 8 && sizes[i] <= stride) {
                const unsigned char *z = streams + (size_t)i * stride;
                std::vector<unsigned char> out;
                out.reserve((size_t)sizes[i] + 128);
                out.insert(out.end(), kSig, kSig + 8);
                unsigned char ihdr[13];
                put32(ihdr, (uint32_t)W); put32(ihdr + 4, (uint32_t)H);
                ihdr[8] = 8; ihdr[9] = 6; ihdr[10] = ihdr[11] = ihdr[12] = 0;
                chunk(&out, "IHDR", ihdr, 13);
                chunk(&out, "IDAT", z, sizes[i]);
                chunk(&out, "IEND", nullptr, 0);
                std::string tmp = std::string(paths[i]) + ".part";
                FILE *f = fopen(tmp.c_str(), "wb");
                if (f) {
                    bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
                    ok = (fclose(f) == 0) && ok;
                    if (ok && rename(tmp.c_str(), paths[i]) == 0) r = RR_OK;
                    else remove(tmp.c_str());
                }
            }
        } catch (...) { r = RR_ERR_ARG; }
        if (r != RR_OK) bad.fetch_add(1);
    });
    if (bad.load()) rr_set_error("rr_host_png_write_streams: some files could not be written");
    return bad.load();
}

// test hook of rr_host_inflate.h: zlib stream -> exactly out_len bytes (RR_ERR_ARG for anything malformed or of another size)
extern "C" int rr_host_zlib_decompress_fast(const uint8_t *z, size_t zlen, uint8_t *out, size_t out_len) {
    if (!z || (!out && out_len)) { rr_set_error("rr_host_zlib_decompress_fast: bad arguments"); return RR_ERR_ARG; }
    try {
        std::vector<unsigned char> buf(z, z + zlen);
        buf.resize(zlen + 16, 0);
        std::vector<unsigned char> tmp(out_len + 8);
        if (!rr_inflate::zlib_decompress(buf.data(), zlen, tmp.data(), out_len)) return RR_ERR_ARG;
        if (out_len) memcpy(out, tmp.data(), out_len);
    } catch (...) { return RR_ERR_ARG; }
    return RR_OK;
}

// test hook of rr_host_deflate.h: data -> zlib stream (any inflate must reproduce data)
extern "C" int rr_host_zlib_compress_fast(const uint8_t *data, size_t n, uint8_t *out, size_t cap, size_t *out_len) {
    if ((n && !data) || !out || !out_len) { rr_set_error("rr_host_zlib_compress_fast: bad arguments"); return RR_ERR_ARG; }
    try {
        std::vector<unsigned char> z;
        rr_deflate::zlib_compress_fast(data, n, z);
        if (z.size() > cap) { rr_set_error("rr_host_zlib_compress_fast: output buffer too small"); return RR_ERR_CAPACITY; }
        memcpy(out, z.data(), z.size());
        *out_len = z.size();
    } catch (...) { rr_set_error("rr_host_zlib_compress_fast: out of memory"); return RR_ERR_ARG; }
    return RR_OK;
}

// Encodes n frames: 8-bit RGB files from bgr (n x H x W x 3, BGR order like the cv2.imwrite input) and 16-bit gray
// mask files from the float32 masks; either list may be NULL.  Returns the number of files that failed.
extern "C" int rr_host_png_write_batch(int n, const char *const *image_paths, const uint8_t *bgr, const char *const *mask_paths,
                                       const float *mask, int W, int H, int level, int n_threads) {
    if (n < 0 || W <= 0 || H <= 0 || (image_paths && !bgr) || (mask_paths && !mask) || level < 0 || level > 9) {
        rr_set_error("rr_host_png_write_batch: bad arguments");
        return RR_ERR_ARG;
    }
    std::atomic<int> bad(0);
    parallel_for(2 * n, n_threads, [&](int j) {
        const int i = j >> 1, kind = j & 1;
        int r = RR_OK;
        try {
            if (kind == 0 && image_paths && image_paths[i]) r = write_bgr8(image_paths[i], bgr + (size_t)i * 3 * W * H, W, H, level);
            if (kind == 1 && mask_paths && mask_paths[i]) r = write_mask16(mask_paths[i], mask + (size_t)i * W * H, W, H, level);
        } catch (...) { r = RR_ERR_ARG; }
        if (r != RR_OK) bad.fetch_add(1);
    });
    if (bad.load()) rr_set_error("rr_host_png_write_batch: some files could not be written");
    return bad.load();
}
