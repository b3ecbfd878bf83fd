// A single-shot zlib-stream decoder for PNG image data (host code, no dependency).
//
// Decoding a frame's two PNGs is what bounds the drop-in pipeline once rendering takes 0.1 ms per frame, and 80 % of
// that is zlib's inflate (a byte-oriented state machine built for streaming).  PNG gives us the whole compressed stream
// and the exact decoded size up front, so this decoder works in one pass over contiguous buffers: a 64-bit bit buffer
// refilled eight bytes at a time, a 12-bit first-level table for the literal/length code (literal pairs where two codes fit) and an 8-bit one for the
// distance code (second-level tables behind them for longer codes), literals decoded back to back, matches copied in
// 8-byte steps.  Every write is bounds-checked against the known output size and every malformed code is an error:
// corrupt input yields RR_ERR_ARG, never an overrun.  The Adler-32 trailer is verified.
//
// RFC 1950 (zlib container), RFC 1951 (deflate: stored, fixed and dynamic Huffman blocks).
#pragma once
#include <stdint.h>
#include <string.h>
#include <zlib.h>

namespace rr_inflate {

enum { LIT_BITS = 12, DIST_BITS = 8, LIT_TABLE = (1 << LIT_BITS) + 288 * 16, DIST_TABLE = (1 << DIST_BITS) + 32 * 128 };
// table entry: low 8 bits = code bits to consume (0 = invalid code); bits 8..11 = extra bits (lengths / distances), the
// width of the second-level table (KIND_SUB) or the number of literals the entry carries (KIND_LIT: 1 or 2); bits 12..15
// kind; bits 16..31 = literal value(s) (first in the low byte) / base / sub-table offset.
// Literal PAIRS: PNG image data that barely compresses is almost all literals with short codes (5 - 6 bits on average), so
// two consecutive literals often fit into one first-level look-up; pair_literals() rewrites those entries after the table is
// built, and the decoder stores 16 bits and advances by the entry's count.
enum { KIND_LIT = 1, KIND_LEN = 2, KIND_EOB = 3, KIND_SUB = 4, KIND_DIST = 5 };
static inline uint32_t entry(int kind, int bits, int extra, int value) { return (uint32_t)bits | ((uint32_t)extra << 8) | ((uint32_t)kind << 12) | ((uint32_t)value << 16); }

static const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

// Builds the two-level decode table of a canonical Huffman code.  lens[0..n): code lengths (0 = unused).  what: 0 literal /
// length alphabet, 1 distance alphabet, 2 code-length alphabet (values = the symbol itself).  Returns false when the code
// is over-subscribed, or incomplete (except the single-code case deflate allows).
//
// A deflate block of OpenCV's / libpng's writers holds about 16 K symbols, so a 1.4 MB image rebuilds its tables 85 times:
// the build is written for speed.  Symbols are visited in canonical order (counting sort by length); the first-level table
// grows with the code length -- while the codes are l bits long only the first 2^l entries exist, and moving on to l + 1
// bits DOUBLES the table with one memcpy (an l-bit code's entry repeats every 2^l indices) -- so every entry is written
// once, sequentially, instead of by one strided loop per symbol.  The stream's bit order is LSB first, so the table is
// indexed by the bit-REVERSED code; `code` below is kept reversed and incremented from its top bit.
static inline uint32_t symbol_entry(int what, int s) {
    if (what == 0) {
        if (s < 256) return entry(KIND_LIT, 0, 1, s);
        if (s == 256) return entry(KIND_EOB, 0, 0, 0);
        if (s <= 285) return entry(KIND_LEN, 0, kLenExtra[s - 257], kLenBase[s - 257]);
        return 0;                                                  // 286, 287: never valid in data
    }
    if (what == 1) return s < 30 ? entry(KIND_DIST, 0, kDistExtra[s], kDistBase[s]) : 0;
    return entry(KIND_LIT, 0, 1, s);
}

static inline bool build_table(const uint8_t *lens, int n, int what, uint32_t *table, int primary_bits, int table_cap) {
    int count[16] = {0};
    for (int i = 0; i < n; i++) count[lens[i]]++;
    count[0] = 0;
    int max_len = 0, min_len = 0, used = 0;
    for (int l = 15; l >= 1; l--) if (count[l]) { if (!max_len) max_len = l; min_len = l; used += count[l]; }
    // Kraft sum
    long left = 1;
    for (int l = 1; l <= 15; l++) { left <<= 1; left -= count[l]; if (left < 0) return false; }
    if (used == 0) {                                               // no distance codes at all: legal when the block has no matches
        for (int i = 0; i < (1 << primary_bits); i++) table[i] = 0;
        return what == 1;
    }
    if (left > 0) {                                                // incomplete: only a lone 1-bit code is allowed
        if (!(used == 1 && max_len == 1)) return false;
        int s = 0;
        while (!lens[s]) s++;
        uint32_t e = symbol_entry(what, s);
        if (e) e |= 1u;
        for (int i = 0; i < (1 << primary_bits); i += 2) { table[i] = e; table[i + 1] = 0; }
        return true;
    }
    // symbols in canonical order
    uint16_t sorted[320];
    int offs[17];
    offs[1] = 0;
    for (int l = 1; l < 16; l++) offs[l + 1] = offs[l] + count[l];
    for (int s = 0; s < n; s++) if (lens[s]) sorted[offs[lens[s]]++] = (uint16_t)s;
    // ---- first level ----
    uint32_t code = 0;                                             // bit-reversed code of the next symbol
    int k = 0, len = min_len;
    int cur_bits = len < primary_bits ? len : primary_bits;        // the table holds 2^cur_bits entries so far
    for (; len <= primary_bits && len <= max_len; len++) {
        for (; cur_bits < len; cur_bits++) memcpy(table + (1u << cur_bits), table, sizeof(uint32_t) << cur_bits);
        const uint32_t ones = (1u << len) - 1;
        for (int c = count[len]; c > 0; c--) {
            uint32_t e = symbol_entry(what, sorted[k++]);
            if (e) e |= (uint32_t)len;
            table[code] = e;
            if (code == ones) {                                    // the all-ones code: the last one of a complete code
                for (; cur_bits < primary_bits; cur_bits++) memcpy(table + (1u << cur_bits), table, sizeof(uint32_t) << cur_bits);
                return k == used;
            }
            const uint32_t bit = 1u << (31 - __builtin_clz(code ^ ones));     // highest 0 bit: the carry of the increment stops there
            code = (code & (bit - 1)) | bit;
        }
    }
    for (; cur_bits < primary_bits; cur_bits++) memcpy(table + (1u << cur_bits), table, sizeof(uint32_t) << cur_bits);
    // ---- codes longer than the first level: one second-level table of 2^(max_len - primary_bits) entries per prefix; codes
    // with the same first `primary_bits` stream bits are neighbours in canonical order ----
    const int sub_bits = max_len - primary_bits;
    const uint32_t pmask = (1u << primary_bits) - 1;
    int next_sub = 1 << primary_bits;
    uint32_t cur_prefix = 0xffffffffu, base = 0;
    for (; len <= max_len; len++) {
        const uint32_t ones = (1u << len) - 1;
        for (int c = count[len]; c > 0; c--) {
            uint32_t e = symbol_entry(what, sorted[k++]);
            if (e) e |= (uint32_t)(len - primary_bits);
            const uint32_t prefix = code & pmask;
            if (prefix != cur_prefix) {
                if (next_sub + (1 << sub_bits) > table_cap) return false;
                cur_prefix = prefix;
                base = (uint32_t)next_sub;
                next_sub += 1 << sub_bits;
                table[prefix] = entry(KIND_SUB, primary_bits, sub_bits, (int)base);
            }
            for (uint32_t i = code >> primary_bits; i < (1u << sub_bits); i += 1u << (len - primary_bits)) table[base + i] = e;
            if (code == ones) return k == used;
            const uint32_t bit = 1u << (31 - __builtin_clz(code ^ ones));
            code = (code & (bit - 1)) | bit;
        }
    }
    return false;                                                  // unreachable for a complete code (the all-ones code ends it)
}

// First-level entries whose index holds TWO complete literal codes become pair entries (count 2, both bytes, the bits of
// both codes).  The second code is looked up in a copy of the single-literal table: its entry is replicated over the unknown
// high bits, so index >> l1 finds it whenever l1 + l2 <= primary_bits.
static inline void pair_literals(uint32_t *table, int primary_bits) {
    static_assert(LIT_BITS <= 12, "pair_literals' copy holds 4096 entries");
    const int n = 1 << primary_bits;
    uint32_t single[1 << 12];
    memcpy(single, table, sizeof(uint32_t) * n);
    // an l1-bit literal owns the entries i0 + (k << l1), i0 < 2^l1: visit every literal once, at its first entry, and walk
    // its copies with k -- single[k] is then the entry of the bits that follow the first code
    for (int i0 = 0; i0 < n / 2; i0++) {
        const uint32_t e = single[i0];
        if (((e >> 12) & 15) != KIND_LIT) continue;
        const int l1 = (int)(e & 0xff);
        if (i0 >> l1) continue;                                     // a copy, not the first entry (also skips l1 = primary_bits)
        const int room = primary_bits - l1;                         // bits left for the second code
        const uint32_t lit1 = (e >> 16) & 0xff;
        for (int k = 0; k < (1 << room); k++) {
            const uint32_t e2 = single[k];
            if (((e2 >> 12) & 15) != KIND_LIT) continue;
            const int l2 = (int)(e2 & 0xff);
            if (l2 > room) continue;
            table[i0 + (k << l1)] = entry(KIND_LIT, l1 + l2, 2, (int)(lit1 | (((e2 >> 16) & 0xff) << 8)));
        }
    }
}

// Adler-32 (RFC 1950) in blocks short enough for 32-bit sums, written so that the compiler vectorises the two sums (an AVX2
// clone is picked at load time where the CPU has it): 0.15 ms for a 1.4 MB image against 0.7 ms in zlib's byte loop.
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
__attribute__((target_clones("avx2", "default")))
#endif
inline uint32_t adler32_blocks(const uint8_t *p, size_t n) {
    uint32_t a = 1, b = 0;
    while (n) {
        const uint32_t L = n < 5552 ? (uint32_t)n : 5552u;        // 255 * L (L + 1) / 2 < 2^32
        uint32_t s1 = 0, s2 = 0;
        for (uint32_t i = 0; i < L; i++) { s1 += p[i]; s2 += (L - i) * p[i]; }
        b = (uint32_t)((b + (uint64_t)L * a + s2) % 65521u);
        a = (a + s1) % 65521u;
        p += L; n -= L;
    }
    return (b << 16) | a;
}

struct Bits {
    const uint8_t *in, *in_end;     // in_end: end of the real data; the buffer carries >= 8 readable bytes beyond it
    uint64_t buf = 0;
    int cnt = 0;
    inline void refill() {          // afterwards cnt >= 56
        uint64_t w;
        memcpy(&w, in, 8);          // little-endian hosts (x86-64 / aarch64)
        buf |= w << cnt;
        in += (63 - cnt) >> 3;
        cnt |= 56;
    }
    inline uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
    inline void drop(int n) { buf >>= n; cnt -= n; }
    // true when more bits have been consumed than the real input holds
    inline bool overrun() const { return in - (cnt >> 3) > in_end; }
};

// in[0..in_len) must be followed by at least 8 readable bytes (any value).  Decodes exactly `out_len` bytes; a stream that
// decodes to anything else is an error.  Returns true on success.
static inline bool zlib_decompress(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len) {
    if (in_len < 6) return false;
    if ((in[0] & 0x0f) != 8 || (in[0] >> 4) > 7 || (((unsigned)in[0] << 8) | in[1]) % 31 != 0 || (in[1] & 0x20)) return false;
    Bits b;
    b.in = in + 2;
    b.in_end = in + in_len - 4;                                    // the Adler-32 is not deflate data
    uint8_t *o = out, *const o_end = out + out_len;
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint32_t lit_table[LIT_TABLE], dist_table[DIST_TABLE], cl_table[1 << 7];
    bool final_block = false;
    while (!final_block) {
        b.refill();
        final_block = b.peek(1);
        const int type = (b.peek(3) >> 1);
        b.drop(3);
        if (type == 0) {                                           // stored
            b.drop(b.cnt & 7);
            b.refill();
            const uint32_t len = b.peek(16), nlen = (uint32_t)((b.buf >> 16) & 0xffff);
            b.drop(32);
            if ((len ^ 0xffff) != nlen) return false;
            // give the unread whole bytes of the bit buffer back to the byte stream
            const uint8_t *src = b.in - (b.cnt >> 3);
            if (src + len > b.in_end || len > (size_t)(o_end - o)) return false;
            memcpy(o, src, len);
            o += len;
            b.in = src + len; b.buf = 0; b.cnt = 0;
            continue;
        }
        if (type == 3) return false;
        if (type == 1) {                                           // fixed Huffman codes
            uint8_t lens[288 + 32];
            for (int i = 0; i < 144; i++) lens[i] = 8;
            for (int i = 144; i < 256; i++) lens[i] = 9;
            for (int i = 256; i < 280; i++) lens[i] = 7;
            for (int i = 280; i < 288; i++) lens[i] = 8;
            for (int i = 0; i < 32; i++) lens[288 + i] = 5;
            if (!build_table(lens, 288, 0, lit_table, LIT_BITS, LIT_TABLE) || !build_table(lens + 288, 32, 1, dist_table, DIST_BITS, DIST_TABLE)) return false;
            pair_literals(lit_table, LIT_BITS);
        } else {                                                   // dynamic Huffman codes
            const int hlit = (int)b.peek(5) + 257, hdist = (int)((b.buf >> 5) & 31) + 1, hclen = (int)((b.buf >> 10) & 15) + 4;
            b.drop(14);
            if (hlit > 286 || hdist > 30) return false;
            uint8_t cl[19] = {0};
            for (int i = 0; i < hclen; i++) {
                if (b.cnt < 3) b.refill();
                cl[order[i]] = (uint8_t)b.peek(3);
                b.drop(3);
            }
            if (!build_table(cl, 19, 2, cl_table, 7, 1 << 7)) return false;
            uint8_t lens[286 + 30 + 140];
            int n = 0;
            while (n < hlit + hdist) {
                b.refill();
                const uint32_t e = cl_table[b.peek(7)];
                if (!(e & 0xff)) return false;
                b.drop((int)(e & 0xff));
                const int sym = (int)(e >> 16);
                if (sym < 16) { lens[n++] = (uint8_t)sym; continue; }
                int rep, val = 0;
                if (sym == 16) { if (n == 0) return false; val = lens[n - 1]; rep = 3 + (int)b.peek(2); b.drop(2); }
                else if (sym == 17) { rep = 3 + (int)b.peek(3); b.drop(3); }
                else { rep = 11 + (int)b.peek(7); b.drop(7); }
                if (n + rep > hlit + hdist) return false;
                while (rep--) lens[n++] = (uint8_t)val;
            }
            if (b.overrun() || lens[256] == 0) return false;       // a block without an end-of-block code cannot end
            if (!build_table(lens, hlit, 0, lit_table, LIT_BITS, LIT_TABLE) || !build_table(lens + hlit, hdist, 1, dist_table, DIST_BITS, DIST_TABLE)) return false;
            pair_literals(lit_table, LIT_BITS);
        }
        // ---- the block's symbols ----
        uint8_t *const o_fast_end = out_len > 320 ? o_end - 320 : out;     // below it four literals and a match need no bound checks
        for (;;) {
            b.refill();
            if (b.in > b.in_end && b.overrun()) return false;
            uint32_t e = lit_table[b.peek(LIT_BITS)];
            if (o < o_fast_end) {
                // fast path: up to four look-ups (eight literals) per refill: 4 x 12 bits of first-level codes fit the 56 bits a refill guarantees.  A literal entry stores two bytes and advances by its count (the second byte of a single
                // is overwritten by whatever comes next; the output has 320 bytes of head-room here).
#define RR_INF_LIT() { b.drop((int)(e & 0xff)); const uint16_t v_ = (uint16_t)(e >> 16); memcpy(o, &v_, 2); o += (e >> 8) & 15; e = lit_table[b.peek(LIT_BITS)]; }
                if (((e >> 12) & 15) == KIND_LIT) {
                    RR_INF_LIT()
                    if (((e >> 12) & 15) == KIND_LIT) {
                        RR_INF_LIT()
                        if (((e >> 12) & 15) == KIND_LIT) {
                            RR_INF_LIT()
                            if (((e >> 12) & 15) == KIND_LIT) {
                                b.drop((int)(e & 0xff)); const uint16_t v_ = (uint16_t)(e >> 16); memcpy(o, &v_, 2); o += (e >> 8) & 15;
                                continue;
                            }
                        }
                    }
                }
#undef RR_INF_LIT
            } else {
                // near the end of the output: one entry at a time, every write checked
                int guard = 0;
                while (((e >> 12) & 15) == KIND_LIT && guard < 3) {
                    const int c = (int)((e >> 8) & 15);
                    if ((size_t)(o_end - o) < (size_t)c) return false;
                    b.drop((int)(e & 0xff));
                    o[0] = (uint8_t)(e >> 16);
                    if (c == 2) o[1] = (uint8_t)(e >> 24);
                    o += c;
                    e = lit_table[b.peek(LIT_BITS)];
                    guard++;
                }
                if (guard == 3) continue;                          // refill before going on
            }
            int kind = (int)((e >> 12) & 15);
            if (kind == KIND_SUB) {
                b.drop((int)(e & 0xff));
                e = lit_table[(e >> 16) + b.peek((int)((e >> 8) & 15))];
                kind = (int)((e >> 12) & 15);
                if (b.cnt < 40) b.refill();
            }
            if (!(e & 0xff)) return false;                         // invalid code
            b.drop((int)(e & 0xff));
            if (kind == KIND_LIT) {                                // a second-level literal (single), or a first-level entry the paths above left
                const int c = (int)((e >> 8) & 15);
                if ((size_t)(o_end - o) < (size_t)c) return false;
                o[0] = (uint8_t)(e >> 16);
                if (c == 2) o[1] = (uint8_t)(e >> 24);
                o += c;
                continue;
            }
            if (kind == KIND_EOB) break;
            if (kind != KIND_LEN) return false;
            const int xb = (int)((e >> 8) & 15);
            const size_t len = (size_t)(e >> 16) + b.peek(xb);
            b.drop(xb);
            if (b.cnt < 32) b.refill();
            uint32_t d = dist_table[b.peek(DIST_BITS)];
            if (((d >> 12) & 15) == KIND_SUB) {
                b.drop((int)(d & 0xff));
                d = dist_table[(d >> 16) + b.peek((int)((d >> 8) & 15))];
            }
            if (!(d & 0xff) || ((d >> 12) & 15) != KIND_DIST) return false;
            b.drop((int)(d & 0xff));
            const int dxb = (int)((d >> 8) & 15);
            const size_t dist = (size_t)(d >> 16) + b.peek(dxb);
            b.drop(dxb);
            if (dist > (size_t)(o - out) || len > (size_t)(o_end - o)) return false;
            const uint8_t *s = o - dist;
            if (dist >= 8 && (size_t)(o_end - o) >= len + 8) {     // 8 bytes at a time (may write up to 7 bytes past the match, inside the buffer)
                uint8_t *q = o;
                const uint8_t *const qe = o + len;
                do { uint64_t w; memcpy(&w, s, 8); memcpy(q, &w, 8); s += 8; q += 8; } while (q < qe);
            } else if (dist == 1) memset(o, *s, len);
            else if (len >= 24 && (size_t)(o_end - o) >= len + 8) {
                // a short period (2 .. 7): 8 * dist bytes of the repeating sequence are dist whole 64-bit words; they are built
                // once in registers / the stack and stored round robin -- no load ever touches bytes just written (a copy from
                // "dist rounded up to 8" bytes back stalls on store forwarding whenever that is not a multiple of 8)
                uint8_t pat[56];
                for (size_t i = 0, k = 0; i < 8 * dist; i++) { pat[i] = s[k]; if (++k == dist) k = 0; }
                uint64_t pw[7];
                memcpy(pw, pat, 8 * dist);
                uint8_t *q = o;
                const uint8_t *const qe = o + len;
                for (size_t k = 0; q < qe; q += 8) { memcpy(q, &pw[k], 8); if (++k == dist) k = 0; }
            } else for (size_t i = 0; i < len; i++) o[i] = s[i];
            o += len;
        }
        if (b.overrun()) return false;
    }
    if (o != o_end) return false;
    // Adler-32 of the decoded data, big-endian, right after the last (byte-aligned) deflate byte
    const uint8_t *tail = b.in - (b.cnt >> 3);
    if (tail > b.in_end) return false;
    const uint8_t *ad = in + in_len - 4;
    const uint32_t want = ((uint32_t)ad[0] << 24) | ((uint32_t)ad[1] << 16) | ((uint32_t)ad[2] << 8) | ad[3];
    return adler32_blocks(out, out_len) == want;
}

}  // namespace rr_inflate
