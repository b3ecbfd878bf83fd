// PNG image data produced on the GPU: what Generator.run saves (common/generator.py:466-467, plt.imsave -> 8-bit RGBA
// files) leaves the device as finished zlib streams instead of pixels.
//
// The host then only frames each stream (signature, IHDR, IDAT length + CRC-32, IEND -- csrc/rr_host_png.cpp) and writes
// the file: the Sub filter, the Adler-32 and the whole deflate step -- 60 % of the drop-in pipeline's CPU time per frame
// once the renderer takes 0.1 ms -- run here, and fewer bytes cross the host link than the pixels would take.
//
// Per frame and file (rainy image: RGBA from the uint8 BGR output; rain mask: matplotlib's viridis of the colormap index):
//   k_png_filter   Sub-filtered scanlines (filter byte + 4 bytes per pixel) -> `filt`; byte histogram; Adler-32 sums
//   k_png_codes    length-limited (15 bit) Huffman code of the histogram, canonical codes, the dynamic-block header, the
//                  stream size and the Adler-32 trailer -- one block per stream, the serial part on one thread
//   k_png_chunk_bits + k_png_scan   bit offset of every 128-byte chunk of the filtered stream
//   k_png_emit     every thread packs its chunk's codes; words shared with a neighbour are merged with atomicOr
// One dynamic-Huffman block per stream, literals only (RFC 1951: HDIST = 0 with a zero-length distance code says so).
// Deterministic: integer atomics only, and the emitted bits do not depend on their order.
#include <cuda_runtime.h>
#include <stdint.h>
#include "rr_png_gpu.cuh"
#include "rr_viridis.h"

__constant__ uint8_t c_viridis[256 * 3];

cudaError_t rr_png_upload_constants() { return cudaMemcpyToSymbol(c_viridis, rr_viridis_rgb, sizeof(rr_viridis_rgb)); }

#define PNG_CHUNK 128                 // source bytes per emitting thread
#define PNG_HEADER_BITS 1106          // 3 + 5 + 5 + 4 + 19 * 3 + 258 * 4
#define PNG_DATA_BIT0 (16 + PNG_HEADER_BITS)      // the zlib header's two bytes come first

// ---- Sub filter + histogram + Adler-32 sums: one block per (scanline, stream) -------------------------------------------
template <bool MASK>
__global__ void __launch_bounds__(256) k_png_filter(const uint8_t *src, rr_png_bufs p) {
    extern __shared__ __align__(16) unsigned char s_row[];          // [misalignment + 1 + 4 W], then the histogram
    const int y = blockIdx.x, f = blockIdx.y, tid = threadIdx.x;
    const int W = p.W;
    const size_t row_bytes = (size_t)4 * W + 1, r0 = (size_t)y * row_bytes;
    unsigned *hist = (unsigned *)(s_row + ((row_bytes + 3 + 15) & ~(size_t)15));
    for (int i = tid; i < 256; i += 256) hist[i] = 0;
    const int mis = (int)(r0 & 3);                                  // the row's first byte inside its 32-bit word
    __syncthreads();
    unsigned char *row = s_row + mis;
    unsigned long long s1 = 0, s2 = 0;
    const unsigned long long n = p.n;
    if (tid == 0) { row[0] = 1; atomicAdd(&hist[1], 1u); s1 += 1; s2 += (n - r0); }      // filter type Sub
    const uint8_t *img = src + (size_t)f * W * p.H * (MASK ? 1 : 3) + (size_t)y * W * (MASK ? 1 : 3);
    for (int x = tid; x < W; x += 256) {
        unsigned r, g, b, pr = 0, pg = 0, pb = 0;
        if (MASK) {
            const uint8_t *c = c_viridis + 3 * img[x];
            r = c[0]; g = c[1]; b = c[2];
            if (x > 0) { const uint8_t *q = c_viridis + 3 * img[x - 1]; pr = q[0]; pg = q[1]; pb = q[2]; }
        } else {
            r = img[3 * x + 2]; g = img[3 * x + 1]; b = img[3 * x];
            if (x > 0) { pr = img[3 * x - 1]; pg = img[3 * x - 2]; pb = img[3 * x - 3]; }
        }
        const unsigned d0 = (r - pr) & 255u, d1 = (g - pg) & 255u, d2 = (b - pb) & 255u, d3 = x == 0 ? 255u : 0u;
        unsigned char *o = row + 1 + 4 * x;
        o[0] = (unsigned char)d0; o[1] = (unsigned char)d1; o[2] = (unsigned char)d2; o[3] = (unsigned char)d3;
        atomicAdd(&hist[d0], 1u); atomicAdd(&hist[d1], 1u); atomicAdd(&hist[d2], 1u); atomicAdd(&hist[d3], 1u);
        const unsigned long long i0 = r0 + 1 + (size_t)4 * x;
        s1 += d0 + d1 + d2 + d3;
        s2 += (n - i0) * d0 + (n - i0 - 1) * d1 + (n - i0 - 2) * d2 + (n - i0 - 3) * d3;
    }
    __syncthreads();
    // the row leaves as aligned 32-bit words; the words it shares with its neighbours byte by byte
    uint8_t *dst = p.filt + (size_t)f * p.n_pad;
    const size_t a0 = r0 - mis, a1 = r0 + row_bytes;               // [a0, a1) covers the row, a0 word aligned
    const size_t nwords = (a1 - a0 + 3) >> 2;
    for (size_t w = tid; w < nwords; w += 256) {
        const size_t g0 = a0 + 4 * w;
        if (g0 >= r0 && g0 + 4 <= a1) *(unsigned *)(dst + g0) = *(const unsigned *)(s_row + 4 * w);
        else for (int k = 0; k < 4; k++) if (g0 + k >= r0 && g0 + k < a1) dst[g0 + k] = s_row[4 * w + k];
    }
    for (int i = tid; i < 256; i += 256) if (hist[i]) atomicAdd(&p.hist[(size_t)f * 256 + i], hist[i]);
    // block sums of the Adler terms
    __shared__ unsigned long long red[2][8];
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if ((tid & 31) == 0) { red[0][tid >> 5] = s1; red[1][tid >> 5] = s2; }
    __syncthreads();
    if (tid == 0) {
        unsigned long long t1 = 0, t2 = 0;
        for (int k = 0; k < 8; k++) { t1 += red[0][k]; t2 += red[1][k]; }
        atomicAdd(&p.adler[(size_t)f * 2], t1);
        atomicAdd(&p.adler[(size_t)f * 2 + 1], t2);
    }
}

// ---- the Huffman code of one stream -------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned png_bitrev(unsigned v, int n) { return __brev(v) >> (32 - n); }

struct png_bits {                      // serial bit writer of the block header (one thread)
    unsigned *out;
    unsigned long long acc;
    int nb;
    size_t w;
    __device__ void put(unsigned v, int n) {
        acc |= (unsigned long long)v << nb;
        nb += n;
        if (nb >= 32) { out[w++] = (unsigned)acc; acc >>= 32; nb -= 32; }
    }
};

__global__ void __launch_bounds__(256) k_png_codes(rr_png_bufs p) {
    __shared__ unsigned freq[260];       // 257 used; padded to whole 128-bit words: the compiler reads the rank loop's operands four at a time
    __shared__ int sorted[257];
    __shared__ unsigned weight[514];
    __shared__ short parent[514];
    __shared__ unsigned char len[257];
    __shared__ int n_used;
    const int f = blockIdx.x, tid = threadIdx.x;
    freq[tid] = p.hist[(size_t)f * 256 + tid];
    if (tid == 0) { freq[256] = 1; n_used = 0; }                    // the end-of-block symbol occurs once
    __syncthreads();
    // rank sort of the used symbols by (frequency, symbol)
    for (int s = tid; s < 257; s += 256) {
        len[s] = 0;
        if (!freq[s]) continue;
        int rank = 0;
        for (int j = 0; j < 257; j++) rank += freq[j] && (freq[j] < freq[s] || (freq[j] == freq[s] && j < s));
        sorted[rank] = s;
        atomicAdd(&n_used, 1);
    }
    __syncthreads();
    if (tid != 0) return;
    const int m = n_used;
    unsigned *out = (unsigned *)(p.stream + (size_t)f * p.cap);
    if (m == 1) len[sorted[0]] = 1;                                 // only the end-of-block symbol (empty stream): one 1-bit code
    else {
        // two-queue Huffman merge: leaves in rank order, internal nodes appear in non-decreasing weight order
        for (int i = 0; i < m; i++) weight[i] = freq[sorted[i]];
        int li = 0, ii = m, nn = m;
        while ((m - li) + (nn - ii) > 1) {
            int a, b;
            if (li < m && (ii >= nn || weight[li] <= weight[ii])) a = li++; else a = ii++;
            if (li < m && (ii >= nn || weight[li] <= weight[ii])) b = li++; else b = ii++;
            weight[nn] = weight[a] + weight[b];
            parent[a] = (short)nn; parent[b] = (short)nn;
            nn++;
        }
        // depths from the root down (a parent's index is above its children's); weight[] is reused for them
        int count[48];
        for (int d = 0; d < 48; d++) count[d] = 0;
        weight[nn - 1] = 0;
        for (int i = nn - 2; i >= 0; i--) {
            weight[i] = weight[parent[i]] + 1;
            if (i < m) count[weight[i] < 47 ? weight[i] : 47]++;
        }
        int deepest = 0;
        for (int d = 0; d < 48; d++) if (count[d]) deepest = d;
        if (deepest > 15) {                                         // the classic repair of an over-deep code (zlib, miniz)
            for (int d = 16; d < 48; d++) { count[15] += count[d]; count[d] = 0; }
            unsigned long long total = 0;
            for (int d = 1; d <= 15; d++) total += (unsigned long long)count[d] << (15 - d);
            while (total > (1ull << 15)) {
                count[15]--;
                for (int d = 14; d >= 1; d--) if (count[d]) { count[d]--; count[d + 1] += 2; break; }
                total--;
            }
            deepest = 15;
        }
        int k = 0;
        for (int d = deepest; d >= 1; d--) for (int c = 0; c < count[d]; c++) len[sorted[k++]] = (unsigned char)d;   // rarest symbols, longest codes
    }
    // canonical codes (RFC 1951 3.2.2), stored bit-reversed for LSB-first emission: code | len << 16
    int bl[16], next[16];
    for (int i = 0; i < 16; i++) bl[i] = 0;
    for (int s = 0; s < 257; s++) bl[len[s]]++;
    bl[0] = 0;
    int code = 0;
    next[0] = 0;
    for (int b = 1; b < 16; b++) { code = (code + bl[b - 1]) << 1; next[b] = code; }
    unsigned long long bits = 0;
    unsigned *codes = p.codes + (size_t)f * 257;
    for (int s = 0; s < 257; s++) {
        const int l = len[s];
        codes[s] = l ? (png_bitrev((unsigned)next[l]++, l) | ((unsigned)l << 16)) : 0u;
        bits += (unsigned long long)freq[s] * l;
    }
    // zlib header + the dynamic block's header.  Code lengths are sent plainly, one 4-bit code each: the code-length
    // alphabet gives its symbols 0..15 four bits and leaves the repeat symbols 16..18 out.
    png_bits w;
    w.out = out; w.acc = 0; w.nb = 0; w.w = 0;
    w.put(0x0178u, 16);                                             // CMF 0x78, FLG 0x01
    w.put(1u, 1); w.put(2u, 2);                                     // BFINAL, BTYPE = dynamic
    w.put(0u, 5); w.put(0u, 5); w.put(15u, 4);                      // HLIT = 257, HDIST = 1, HCLEN = 19
    for (int i = 0; i < 19; i++) w.put(i < 3 ? 0u : 4u, 3);         // order 16, 17, 18, then 0, 8, 7, ... : all of 0..15 get 4 bits
    for (int s = 0; s < 257; s++) w.put(png_bitrev(len[s], 4), 4);
    w.put(png_bitrev(0u, 4), 4);                                    // the one distance code: length 0, no distance codes in use
    if (w.nb) atomicOr(&out[w.w], (unsigned)w.acc);                 // the data continues in this word
    // size of the stream, and the Adler-32 behind the last deflate byte
    const unsigned long long total_bits = PNG_DATA_BIT0 + bits;
    const size_t deflate_end = (size_t)((total_bits + 7) >> 3);
    const unsigned long long n = p.n;
    const unsigned a1 = (unsigned)((1 + p.adler[(size_t)f * 2]) % 65521ull);
    const unsigned a2 = (unsigned)((n + p.adler[(size_t)f * 2 + 1]) % 65521ull);
    const unsigned ad = (a2 << 16) | a1;
    unsigned size = (unsigned)(deflate_end + 4);
    if (deflate_end + 4 > p.cap) size = 0;                          // cannot happen (cap is 1.25 n); 0 tells the host to fall back
    else {
        uint8_t *bytes = p.stream + (size_t)f * p.cap;
        bytes[deflate_end] = (uint8_t)(ad >> 24); bytes[deflate_end + 1] = (uint8_t)(ad >> 16);
        bytes[deflate_end + 2] = (uint8_t)(ad >> 8); bytes[deflate_end + 3] = (uint8_t)ad;
    }
    p.sizes[f] = size;
}

// ---- bit offsets of the chunks ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_png_chunk_bits(rr_png_bufs p) {
    __shared__ unsigned char lens[256];
    const int f = blockIdx.y, tid = threadIdx.x;
    lens[tid] = (unsigned char)(p.codes[(size_t)f * 257 + tid] >> 16);
    __syncthreads();
    const int c = blockIdx.x * 256 + tid;
    if (c >= p.nchunks) return;
    const size_t b0 = (size_t)c * PNG_CHUNK;
    const int nb = (int)(p.n - b0 < PNG_CHUNK ? p.n - b0 : PNG_CHUNK);
    const uint4 *src = (const uint4 *)(p.filt + (size_t)f * p.n_pad + b0);
    unsigned bits = 0;
    for (int q = 0; q < PNG_CHUNK / 16; q++) {
        if (q * 16 >= nb) break;
        const uint4 v = src[q];
        const unsigned wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 16; k++) if (q * 16 + k < nb) bits += lens[(wv[k >> 2] >> (8 * (k & 3))) & 255u];
    }
    p.chunk_bits[(size_t)f * p.nchunks + c] = bits;
}

__global__ void __launch_bounds__(1024) k_png_scan(rr_png_bufs p) {
    __shared__ unsigned wtot[32];
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned *cb = p.chunk_bits + (size_t)f * p.nchunks;
    const int per = (p.nchunks + 1023) / 1024;
    const int i0 = tid * per, i1 = i0 + per < p.nchunks ? i0 + per : p.nchunks;
    unsigned t = 0;
    for (int i = i0; i < i1; i++) t += cb[i];
    unsigned inc = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
    if (lane == 31) wtot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const unsigned v = wtot[lane];
        unsigned w = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned u = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += u; }
        wtot[lane] = w - v;
    }
    __syncthreads();
    unsigned run = PNG_DATA_BIT0 + wtot[warp] + inc - t;
    for (int i = i0; i < i1; i++) { const unsigned b = cb[i]; cb[i] = run; run += b; }
}

// ---- emission ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_png_emit(rr_png_bufs p) {
    __shared__ unsigned lut[257];
    const int f = blockIdx.y, tid = threadIdx.x;
    for (int i = tid; i < 257; i += 256) lut[i] = p.codes[(size_t)f * 257 + i];
    __syncthreads();
    const int c = blockIdx.x * 256 + tid;
    if (c >= p.nchunks || p.sizes[f] == 0) return;
    const size_t b0 = (size_t)c * PNG_CHUNK;
    const int nb = (int)(p.n - b0 < PNG_CHUNK ? p.n - b0 : PNG_CHUNK);
    const uint4 *src = (const uint4 *)(p.filt + (size_t)f * p.n_pad + b0);
    unsigned *out = (unsigned *)(p.stream + (size_t)f * p.cap);
    const unsigned pos = p.chunk_bits[(size_t)f * p.nchunks + c];
    size_t w = pos >> 5;
    int have = (int)(pos & 31);                                     // bits of word w that belong to the chunk before
    unsigned long long acc = 0;
    bool first = true;                                              // word w is shared with the previous chunk
    for (int q = 0; q < PNG_CHUNK / 16; q++) {
        if (q * 16 >= nb) break;
        const uint4 v = src[q];
        const unsigned wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (q * 16 + k < nb) {
                const unsigned e = lut[(wv[k >> 2] >> (8 * (k & 3))) & 255u];
                acc |= (unsigned long long)(e & 0xffffu) << have;
                have += (int)(e >> 16);
                if (have >= 32) {
                    if (first) atomicOr(&out[w], (unsigned)acc); else out[w] = (unsigned)acc;
                    first = false;
                    w++; acc >>= 32; have -= 32;
                }
            }
        }
    }
    if (c == p.nchunks - 1) {                                       // the end-of-block code closes the stream
        const unsigned e = lut[256];
        acc |= (unsigned long long)(e & 0xffffu) << have;
        have += (int)(e >> 16);
        if (have >= 32) {
            if (first) atomicOr(&out[w], (unsigned)acc); else out[w] = (unsigned)acc;
            first = false;
            w++; acc >>= 32; have -= 32;
        }
    }
    if (have > 0) atomicOr(&out[w], (unsigned)acc);                 // shared with the next chunk (or with the Adler-32 bytes)
}

cudaError_t rr_launch_png_encode(const rr_png_bufs &p, const uint8_t *src, bool mask, int F, cudaStream_t st) {
    cudaError_t e;
    // histogram, Adler sums and the stream words start from zero (the emitters OR into shared words)
    if ((e = cudaMemsetAsync(p.hist, 0, sizeof(unsigned) * 256 * F, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(p.adler, 0, sizeof(unsigned long long) * 2 * F, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(p.stream, 0, p.cap * F, st)) != cudaSuccess) return e;
    const size_t row_bytes = (size_t)4 * p.W + 1;
    const size_t smem = ((row_bytes + 3 + 15) & ~(size_t)15) + 256 * sizeof(unsigned);
    dim3 g0(p.H, F);
    if (mask) k_png_filter<true><<<g0, 256, smem, st>>>(src, p);
    else k_png_filter<false><<<g0, 256, smem, st>>>(src, p);
    k_png_codes<<<F, 256, 0, st>>>(p);
    dim3 g1((p.nchunks + 255) / 256, F);
    k_png_chunk_bits<<<g1, 256, 0, st>>>(p);
    k_png_scan<<<F, 1024, 0, st>>>(p);
    k_png_emit<<<g1, 256, 0, st>>>(p);
    return cudaGetLastError();
}
