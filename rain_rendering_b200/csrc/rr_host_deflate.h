// A small, fast zlib-stream encoder for filtered PNG scanlines (host code, no dependencies beyond adler32()).
//
// What the frame pipeline needs from deflate is throughput: the renderer delivers thousands of frames per second and
// every frame becomes two PNG files (common/generator.py:466-467).  zlib's own fastest settings (level 1 / Z_HUFFMAN_ONLY /
// Z_RLE) run at 70-100 MB/s per core on this data because they go through the general matcher / tally machinery one byte
// at a time.  This encoder does the two things that pay on Sub-filtered image rows and nothing else:
//   * runs of a repeated byte become distance-1 matches (what Z_RLE finds): the rain mask is mostly one colour, so its
//     filtered rows are long runs of zeros;
//   * everything else is a literal under a per-block dynamic Huffman code (what Z_HUFFMAN_ONLY does).
// Output is a standard zlib stream (RFC 1950/1951): CMF/FLG, dynamic-Huffman blocks, Adler-32.  Any inflate decodes it.
#pragma once
#include <stdint.h>
#include <string.h>
#include <zlib.h>
#include <algorithm>
#include <vector>

namespace rr_deflate {

struct BitWriter {
    unsigned char *p;                                  // write cursor into a buffer the caller sized for the worst case (+ 8 bytes of slack)
    uint64_t acc = 0;
    int nbits = 0;                                     // < 8 after flush()
    explicit BitWriter(unsigned char *dst) : p(dst) {}
    // add up to 56 - 7 bits between two flushes; code already in stream (LSB-first) order
    inline void add(uint32_t code, int len) { acc |= (uint64_t)code << nbits; nbits += len; }
    // branch-free: store the whole accumulator, advance by the complete bytes (little-endian hosts: x86-64 / aarch64)
    inline void flush() {
        memcpy(p, &acc, 8);
        p += nbits >> 3;
        acc >>= (nbits & ~7);
        nbits &= 7;
    }
    inline void put(uint32_t code, int len) { add(code, len); flush(); }
    inline void align() {
        flush();
        if (nbits > 0) { *p++ = (unsigned char)acc; }
        nbits = 0; acc = 0;
    }
};

inline uint32_t bit_reverse(uint32_t v, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; i++) { r = (r << 1) | (v & 1); v >>= 1; }
    return r;
}

// Code lengths (<= max_len) of a Huffman code for freq[0..n): the optimal tree, then the classic overflow repair
// (move the too-deep leaves up to max_len and pay for it by pushing the cheapest shallower leaf down) when it is deeper
// than max_len.  Symbols with freq 0 get length 0; a single used symbol gets length 1.
inline void huffman_lengths(const uint32_t *freq, int n, int max_len, uint8_t *len) {
    struct Node { uint64_t w; int sym, l, r; };
    std::vector<int> used;
    for (int i = 0; i < n; i++) { len[i] = 0; if (freq[i]) used.push_back(i); }
    if (used.empty()) return;
    if (used.size() == 1) { len[used[0]] = 1; return; }
    std::sort(used.begin(), used.end(), [&](int a, int b) { return freq[a] != freq[b] ? freq[a] < freq[b] : a < b; });
    const int m = (int)used.size();
    std::vector<Node> nodes;
    nodes.reserve(2 * m);
    for (int i = 0; i < m; i++) nodes.push_back({freq[used[i]], used[i], -1, -1});
    // two-queue merge: leaves (sorted) and internal nodes (created in non-decreasing weight order)
    int li = 0, ii = m;
    auto take = [&]() -> int {
        if (li < m && (ii >= (int)nodes.size() || nodes[li].w <= nodes[ii].w)) return li++;
        return ii++;
    };
    while ((m - li) + ((int)nodes.size() - ii) > 1) {
        const int a = take(), b = take();
        nodes.push_back({nodes[a].w + nodes[b].w, -1, a, b});
    }
    // depths, iteratively from the root (the last node)
    std::vector<int> depth(nodes.size(), 0);
    std::vector<int> count(64, 0);
    for (int i = (int)nodes.size() - 1; i >= 0; i--) {
        if (nodes[i].sym < 0) { depth[nodes[i].l] = depth[i] + 1; depth[nodes[i].r] = depth[i] + 1; }
        else count[depth[i] < 63 ? depth[i] : 63]++;
    }
    int deepest = 0;
    for (int d = 0; d < 64; d++) if (count[d]) deepest = d;
    if (deepest > max_len) {
        for (int d = max_len + 1; d < 64; d++) { count[max_len] += count[d]; count[d] = 0; }
        // Kraft sum in units of 2^-max_len must come down to exactly 2^max_len
        uint64_t total = 0;
        for (int d = 1; d <= max_len; d++) total += (uint64_t)count[d] << (max_len - d);
        while (total > ((uint64_t)1 << max_len)) {
            count[max_len]--;
            for (int d = max_len - 1; d >= 1; d--)
                if (count[d]) { count[d]--; count[d + 1] += 2; break; }
            total--;
        }
    }
    // hand the lengths out: the rarest symbols get the longest codes
    int k = 0;
    for (int d = max_len < deepest ? max_len : deepest; d >= 1; d--)
        for (int c = 0; c < count[d]; c++) len[used[k++]] = (uint8_t)d;
}

// canonical codes (RFC 1951 3.2.2), returned bit-reversed for LSB-first output
inline void canonical_codes(const uint8_t *len, int n, uint16_t *code) {
    int bl_count[16] = {0}, next[16];
    for (int i = 0; i < n; i++) bl_count[len[i]]++;
    bl_count[0] = 0;
    int c = 0;
    next[0] = 0;
    for (int b = 1; b < 16; b++) { c = (c + bl_count[b - 1]) << 1; next[b] = c; }
    for (int i = 0; i < n; i++) code[i] = len[i] ? (uint16_t)bit_reverse((uint32_t)next[len[i]]++, len[i]) : 0;
}

struct LenCode { uint16_t sym; uint8_t extra_bits; uint16_t base; };
inline const LenCode *length_table() {           // match length 3..258 -> (symbol, extra bits, base)
    static LenCode tab[259];
    static bool init = false;
    if (!init) {
        static const int base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
        static const int ebits[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
        for (int L = 3; L <= 258; L++) {
            int s = 28;
            while (base[s] > L) s--;
            if (L == 258) s = 28;
            tab[L] = {(uint16_t)(257 + s), (uint8_t)ebits[s], (uint16_t)base[s]};
        }
        init = true;
    }
    return tab;
}

// tokens: 0..255 literal, 256 + (len - 3) a distance-1 match of len 3..258; freq: their lit/len symbol counts
inline void write_block(BitWriter &bw, const uint16_t *tok, size_t ntok, uint32_t *freq, bool final_block) {
    const LenCode *LT = length_table();
    freq[256] = 1;
    uint8_t llen[286];
    uint16_t lcode[286];
    huffman_lengths(freq, 286, 15, llen);
    int hlit = 286;
    while (hlit > 257 && llen[hlit - 1] == 0) hlit--;
    canonical_codes(llen, hlit, lcode);
    // one distance code (distance 1 = symbol 0) of one bit; "one distance code ... one unused code" (RFC 1951 3.2.7)
    // ---- block header ----
    bw.put(final_block ? 1u : 0u, 1);
    bw.put(2u, 2);                                       // BTYPE = 10: dynamic Huffman
    bw.put((uint32_t)(hlit - 257), 5);
    bw.put(0u, 5);                                       // HDIST: 1 distance code
    // code-length alphabet: run-length coded lengths (symbols 16 / 17 / 18) under their own Huffman code
    uint8_t seq[286 + 1];
    int nseq = 0;
    for (int i = 0; i < hlit; i++) seq[nseq++] = llen[i];
    seq[nseq++] = 1;                                     // the distance code
    struct CL { uint8_t sym, extra; };
    std::vector<CL> cl;
    for (int i = 0; i < nseq;) {
        int j = i;
        while (j < nseq && seq[j] == seq[i]) j++;
        int run = j - i;
        const uint8_t v = seq[i];
        if (v == 0) {
            while (run >= 11) { const int r = run > 138 ? 138 : run; cl.push_back({18, (uint8_t)(r - 11)}); run -= r; }
            if (run >= 3) { cl.push_back({17, (uint8_t)(run - 3)}); run = 0; }
            while (run-- > 0) cl.push_back({0, 0});
        } else {
            cl.push_back({v, 0});
            run--;
            while (run >= 3) { const int r = run > 6 ? 6 : run; cl.push_back({16, (uint8_t)(r - 3)}); run -= r; }
            while (run-- > 0) cl.push_back({v, 0});
        }
        i = j;
    }
    uint32_t cfreq[19];
    memset(cfreq, 0, sizeof(cfreq));
    for (const CL &c : cl) cfreq[c.sym]++;
    {   // inflate rejects an incomplete code-length code: at least two of its symbols must be in use
        int used = 0;
        for (int i = 0; i < 19; i++) used += cfreq[i] != 0;
        if (used < 2) cfreq[cfreq[0] ? 1 : 0] = 1;
    }
    uint8_t clen[19];
    uint16_t ccode[19];
    huffman_lengths(cfreq, 19, 7, clen);
    canonical_codes(clen, 19, ccode);
    static const int order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    int hclen = 19;
    while (hclen > 4 && clen[order[hclen - 1]] == 0) hclen--;
    bw.put((uint32_t)(hclen - 4), 4);
    for (int i = 0; i < hclen; i++) bw.put(clen[order[i]], 3);
    for (const CL &c : cl) {
        bw.put(ccode[c.sym], clen[c.sym]);
        if (c.sym == 16) bw.put(c.extra, 2);
        else if (c.sym == 17) bw.put(c.extra, 3);
        else if (c.sym == 18) bw.put(c.extra, 7);
    }
    // ---- data ----
    uint32_t lit[256];                                   // code | len << 16
    for (int i = 0; i < 256; i++) lit[i] = (uint32_t)lcode[i] | ((uint32_t)llen[i] << 16);
    size_t i = 0;
    while (i < ntok) {
        // literals three at a time: 3 x 15 bits + 7 pending fit the accumulator
        while (i + 3 <= ntok && (tok[i] | tok[i + 1] | tok[i + 2]) < 256) {
            const uint32_t a = lit[tok[i]], b = lit[tok[i + 1]], c = lit[tok[i + 2]];
            bw.add(a & 0xffffu, (int)(a >> 16));
            bw.add(b & 0xffffu, (int)(b >> 16));
            bw.add(c & 0xffffu, (int)(c >> 16));
            bw.flush();
            i += 3;
        }
        if (i >= ntok) break;
        const unsigned t = tok[i++];
        if (t < 256) { bw.put(lit[t] & 0xffffu, (int)(lit[t] >> 16)); continue; }
        const int L = (int)t - 256 + 3;
        const LenCode &lc = LT[L];
        bw.add(lcode[lc.sym], llen[lc.sym]);
        if (lc.extra_bits) bw.add((uint32_t)(L - lc.base), lc.extra_bits);
        bw.add(0u, 1);                                   // distance symbol 0 (distance 1), its one-bit code "0", no extra bits
        bw.flush();
    }
    bw.put(lcode[256], llen[256]);
}

// data[0..n) -> a complete zlib stream appended to `out`
inline void zlib_compress_fast(const unsigned char *data, size_t n, std::vector<unsigned char> &out) {
    const size_t BLOCK_TOKENS = 1u << 16;
    const size_t base = out.size();
    // worst case: 15 bits per literal, a ~300-byte header per block of 65536 tokens
    out.resize(base + 2 * n + (n / BLOCK_TOKENS + 2) * 512 + 64);
    unsigned char *dst = out.data() + base;
    dst[0] = 0x78; dst[1] = 0x01;
    BitWriter bw(dst + 2);
    const LenCode *LT = length_table();
    std::vector<uint16_t> tokv(BLOCK_TOKENS + 8);
    uint16_t *tok = tokv.data();
    uint32_t freq[286];
    size_t i = 0;
    bool wrote_final = false;
    while (i < n) {
        size_t nt = 0;
        memset(freq, 0, sizeof(freq));
        if (i == 0) { freq[data[0]]++; tok[nt++] = data[0]; i = 1; }      // the stream's first byte has nothing before it
        while (i < n && nt < BLOCK_TOKENS) {
            const unsigned char c = data[i];
            // a run continues the previous byte (distance 1): the byte before it must equal c, and it must be >= 3 long
            if (data[i - 1] == c && i + 2 < n && data[i + 1] == c && data[i + 2] == c) {
                size_t j = i + 3;
                const size_t lim = (n - i) < 258 ? n : i + 258;
                while (j < lim && data[j] == c) j++;
                const int L = (int)(j - i);
                freq[LT[L].sym]++;
                tok[nt++] = (uint16_t)(256 + L - 3);
                i = j;
            } else {
                freq[c]++;
                tok[nt++] = c;
                i++;
            }
        }
        const bool fin = i >= n;
        write_block(bw, tok, nt, freq, fin);
        wrote_final = fin;
    }
    if (!wrote_final) { memset(freq, 0, sizeof(freq)); write_block(bw, nullptr, 0, freq, true); }      // empty input: one empty final block
    bw.align();
    const uLong ad = adler32(adler32(0L, Z_NULL, 0), data, (uInt)n);
    unsigned char *q = bw.p;
    q[0] = (unsigned char)(ad >> 24); q[1] = (unsigned char)(ad >> 16); q[2] = (unsigned char)(ad >> 8); q[3] = (unsigned char)ad;
    out.resize((size_t)(q + 4 - out.data()));
}

}  // namespace rr_deflate
