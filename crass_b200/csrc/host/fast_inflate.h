// A DEFLATE decoder for the feed path: an ordinary .gz is ONE deflate stream, so one thread inflates it and that
// thread is what a run on a compressed file waits for.  zlib keeps a sliding window and copies its output out of it; this decoder
// writes straight into the (contiguous) output mapping, back-references read from there, symbols are decoded through one table
// look-up from a 64-bit bit buffer (11-bit primary table for literals / lengths, 8-bit for distances, sub-tables behind them).
//
// It is deliberately strict: anything it does not like -- a malformed header, an incomplete code that is actually used, a
// distance before the start, input or output running out, a CRC or length that does not match -- makes it give up and the caller
// reads the archive through zlib instead (parser.cpp: reread_like_kseq), which also decides what a DAMAGED archive yields.  Every
// member's CRC-32 is checked, so a wrong output cannot pass for a right one.
//
// RFC 1951 (DEFLATE) and RFC 1952 (gzip) are the specification; nothing of the reference is involved (it calls gzread).
#pragma once
#include <stdint.h>
#include <string.h>
#include <zlib.h>          // crc32() only

#include <atomic>
#include <chrono>
#include <thread>

namespace fastinf {

struct Tables {
    // entry: value << 16 | kind << 12 | extra << 8 | len
    //   kind 0 literal (value = byte), 1 length (value = base, extra = extra bits), 2 end of block, 3 link to a sub-table
    //   (value = first index, extra = its index bits), 7 unused code (an error if it is ever looked up)
    //   len = bits of the codeword (for a sub-table entry: the bits BEYOND the primary ones)
    uint32_t lit[2048 + 4608];
    uint32_t dist[256 + 3840];
};

enum { kLitBits = 11, kDistBits = 8, kKindShift = 12 };
static const uint32_t kLiteralFlag = 0x8000u;     // set in PRIMARY entries of literals only: one test on the fast path

static const uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
static const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t kDistBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

inline uint32_t reverse_bits(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; ++i) { r = (r << 1) | (code & 1u); code >>= 1; }
    return r;
}

// Canonical Huffman code of `n` symbols with the given lengths into a primary table of `pbits` index bits plus sub-tables.
// payload(sym) gives value / kind / extra of a symbol's entry.  false: over-subscribed code or no room for the sub-tables.
template <class Payload>
inline bool build_table(const uint8_t* lens, int n, int pbits, uint32_t* table, int table_cap, Payload payload, uint32_t primary_literal_flag = 0) {
    int count[16] = {0};
    for (int i = 0; i < n; ++i) count[lens[i]]++;
    count[0] = 0;
    int left = 1;
    for (int l = 1; l <= 15; ++l) { left = (left << 1) - count[l]; if (left < 0) return false; }    // over-subscribed
    uint32_t next_code[16];
    uint32_t code = 0;
    for (int l = 1; l <= 15; ++l) { code = (code + (uint32_t)count[l - 1]) << 1; next_code[l] = code; }
    const uint32_t unused = 7u << kKindShift;
    for (int i = 0; i < (1 << pbits); ++i) table[i] = unused;
    // sub-tables: for every primary prefix that long codes share, the longest such code decides the sub-table's size
    uint8_t sub_bits[2048];
    memset(sub_bits, 0, (size_t)1 << pbits);
    {
        uint32_t nc[16];
        memcpy(nc, next_code, sizeof nc);
        for (int s = 0; s < n; ++s) {
            const int l = lens[s];
            if (!l) continue;
            const uint32_t r = reverse_bits(nc[l]++, l);
            if (l > pbits) { const uint32_t pre = r & ((1u << pbits) - 1); if (l - pbits > sub_bits[pre]) sub_bits[pre] = (uint8_t)(l - pbits); }
        }
    }
    int used = 1 << pbits;
    for (int pre = 0; pre < (1 << pbits); ++pre) {
        if (!sub_bits[pre]) continue;
        const int size = 1 << sub_bits[pre];
        if (used + size > table_cap) return false;
        table[pre] = ((uint32_t)used << 16) | (3u << kKindShift) | ((uint32_t)sub_bits[pre] << 8) | (uint32_t)pbits;
        for (int i = 0; i < size; ++i) table[used + i] = unused;
        used += size;
    }
    for (int s = 0; s < n; ++s) {
        const int l = lens[s];
        if (!l) continue;
        const uint32_t r = reverse_bits(next_code[l]++, l);
        const uint32_t pay = payload(s);
        if (l <= pbits) {
            const uint32_t e = pay | (uint32_t)l | (primary_literal_flag && ((pay >> kKindShift) & 7u) == 0 ? primary_literal_flag : 0u);
            for (uint32_t i = r; i < (1u << pbits); i += 1u << l) table[i] = e;
        } else {
            const uint32_t pre = r & ((1u << pbits) - 1);
            const uint32_t link = table[pre];
            const uint32_t base = link >> 16, sb = (link >> 8) & 15u;
            const uint32_t e = pay | (uint32_t)(l - pbits);
            for (uint32_t i = r >> pbits; i < (1u << sb); i += 1u << (l - pbits)) table[base + i] = e;
        }
    }
    return true;
}

inline uint32_t lit_payload(int s) {
    if (s < 256) return (uint32_t)s << 16;
    if (s == 256) return 2u << kKindShift;
    if (s > 285) return 7u << kKindShift;                                   // 286, 287: never valid
    return ((uint32_t)kLenBase[s - 257] << 16) | (1u << kKindShift) | ((uint32_t)kLenExtra[s - 257] << 8);
}
inline uint32_t dist_payload(int s) {
    if (s > 29) return 7u << kKindShift;
    return ((uint32_t)kDistBase[s] << 16) | ((uint32_t)kDistExtra[s] << 8);
}

// The CRC-32 of a large member is computed by a second thread that follows the decoder through the output (a third of the time
// of a one-thread gunzip would go into it otherwise).
struct FollowCrc {
    const uint8_t* base; size_t start;
    std::atomic<size_t> target, done;
    std::atomic<bool> stop{false};
    uint32_t crc;
    std::thread th;
    FollowCrc(const uint8_t* b, size_t s) : base(b), start(s), target(s), done(s), crc((uint32_t)crc32(0L, Z_NULL, 0)) { th = std::thread([this]() { run(); }); }
    void run() {
        size_t pos = start;
        for (;;) {
            const size_t t = target.load(std::memory_order_acquire);
            if (pos < t) {
                const size_t step = t - pos < ((size_t)1 << 20) ? t - pos : ((size_t)1 << 20);
                crc = (uint32_t)crc32(crc, base + pos, (uInt)step);
                pos += step;
                done.store(pos, std::memory_order_release);
            } else if (stop.load(std::memory_order_acquire)) break;
            else std::this_thread::sleep_for(std::chrono::microseconds(50));
        }
    }
    uint32_t finish(size_t end) {                                          // the CRC of [start, end)
        target.store(end, std::memory_order_release);
        while (done.load(std::memory_order_acquire) < end) std::this_thread::sleep_for(std::chrono::microseconds(20));
        stop.store(true, std::memory_order_release);
        th.join();
        return crc;
    }
    ~FollowCrc() { if (th.joinable()) { stop.store(true, std::memory_order_release); th.join(); } }
};

struct Reader {
    const uint8_t* p; const uint8_t* end;
    uint64_t buf = 0; unsigned n = 0;                                        // n valid bits in buf
    unsigned past = 0;                                                       // bytes of zeros supplied after the end of the input
    inline void refill() {                                                   // at least 56 bits
        if (end - p >= 8) {
            uint64_t w;
            memcpy(&w, p, 8);
            buf |= w << n;
            p += (63 - n) >> 3;
            n |= 56;
        } else {
            while (n <= 56) {
                if (p < end) buf |= (uint64_t)*p++ << n;
                else ++past;                                                 // (zeros; whoever consumes them is found out by `past`)
                n += 8;
            }
        }
    }
    inline uint32_t peek(unsigned k) const { return (uint32_t)(buf & ((1ull << k) - 1)); }
    inline void drop(unsigned k) { buf >>= k; n -= k; }
    inline uint32_t take(unsigned k) { const uint32_t v = peek(k); drop(k); return v; }
};

// The blocks of one deflate stream from the reader's position on (RFC 1951), decoded straight into out[o...]; back-references
// reach down to out[member_out].  Returns 1 after a final block, 0 at the first block boundary at or past stop_bit (a bit
// position in `in`), -1 when the decoder gives up.
inline size_t bit_position(const Reader& r, const uint8_t* in) { return (size_t)(r.p - in) * 8 - r.n; }

template <class Progress>
inline int inflate_blocks(Reader& r, const uint8_t* in, Tables* T, uint8_t* out, size_t& o, size_t out_cap, size_t member_out, size_t stop_bit,
                          Progress progress, FollowCrc*& follow) {
    static const uint8_t kOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    for (;;) {
        r.refill();
        if (stop_bit != (size_t)-1 && r.past == 0 && bit_position(r, in) >= stop_bit) return 0;
        const uint32_t bfinal = r.take(1), btype = r.take(2);
        if (btype == 0) {                                                // stored
            r.drop(r.n & 7);                                             // to a byte boundary
            // give the bytes still in the bit buffer back
            const unsigned back = r.n >> 3;
            if (r.past > back) return -1;
            r.p -= (back - r.past); r.buf = 0; r.n = 0; r.past = 0;
            if (r.end - r.p < 4) return -1;
            const uint32_t len = r.p[0] | ((uint32_t)r.p[1] << 8), nlen = r.p[2] | ((uint32_t)r.p[3] << 8);
            if ((len ^ 0xFFFFu) != nlen) return -1;
            r.p += 4;
            if ((size_t)(r.end - r.p) < len || out_cap - o < (size_t)len + 16) return -1;
            memcpy(out + o, r.p, len);
            r.p += len; o += len;
        } else if (btype == 1 || btype == 2) {
            uint8_t lens[320];
            int nlit, ndist;
            if (btype == 1) {
                for (int i = 0; i < 144; ++i) lens[i] = 8;
                for (int i = 144; i < 256; ++i) lens[i] = 9;
                for (int i = 256; i < 280; ++i) lens[i] = 7;
                for (int i = 280; i < 288; ++i) lens[i] = 8;
                for (int i = 0; i < 32; ++i) lens[288 + i] = 5;
                nlit = 288; ndist = 32;
            } else {
                nlit = (int)r.take(5) + 257; ndist = (int)r.take(5) + 1;
                const int ncl = (int)r.take(4) + 4;
                if (nlit > 286 || ndist > 30) return -1;
                uint8_t cl[19];
                memset(cl, 0, sizeof cl);
                for (int i = 0; i < ncl; ++i) { if (r.n < 3) r.refill(); cl[kOrder[i]] = (uint8_t)r.take(3); }
                uint32_t pre[128 + 64];
                if (!build_table(cl, 19, 7, pre, 128 + 64, [](int s) { return (uint32_t)s << 16; })) return -1;
                int i = 0;
                while (i < nlit + ndist) {
                    r.refill();
                    const uint32_t e = pre[r.peek(7)];
                    if (((e >> kKindShift) & 7u) != 0) return -1;   // (code lengths are at most 7 bits: no sub-tables, no unused codes in use)
                    r.drop(e & 15u);
                    const int s = (int)(e >> 16);
                    if (s < 16) { lens[i++] = (uint8_t)s; continue; }
                    int rep; uint8_t v = 0;
                    if (s == 16) { if (!i) return -1; v = lens[i - 1]; rep = 3 + (int)r.take(2); }
                    else if (s == 17) rep = 3 + (int)r.take(3);
                    else rep = 11 + (int)r.take(7);
                    if (i + rep > nlit + ndist) return -1;
                    while (rep--) lens[i++] = v;
                }
                if (!lens[256]) return -1;                           // no end-of-block code
                memmove(lens + 288, lens + nlit, (size_t)ndist);      // (distance lengths to a fixed place)
                memset(lens + nlit, 0, (size_t)(288 - nlit));
            }
            if (!build_table(lens, 288, kLitBits, T->lit, 2048 + 4608, lit_payload, kLiteralFlag)) return -1;
            if (!build_table(lens + 288, btype == 1 ? 32 : ndist, kDistBits, T->dist, 256 + 3840, dist_payload)) return -1;
            // ---- the symbols of the block: first with everything in local variables while at least 16 bytes of input and 320
            //      bytes of room are left (no end-of-input bookkeeping in here), then -- for the tail -- with every check
            bool block_done = false;
            {
                uint64_t bb = r.buf; unsigned bn = r.n;
                const uint8_t* ip = r.p;
                uint8_t* op = out + o;
                uint8_t* const window0 = out + member_out;
                const uint8_t* const ip_safe = (r.end - r.p > 16) ? r.end - 16 : r.p;
                uint8_t* const op_safe = out_cap - o > 320 ? out + out_cap - 320 : op;
                const uint32_t* const lit = T->lit;
                const uint32_t* const dtab = T->dist;
#define FASTINF_REFILL() do { uint64_t w_; memcpy(&w_, ip, 8); bb |= w_ << bn; ip += (63 - bn) >> 3; bn |= 56; } while (0)
#define FASTINF_DROP(k_) do { const unsigned d_ = (k_); bb >>= d_; bn -= d_; } while (0)
                while (ip < ip_safe && op < op_safe) {
                    FASTINF_REFILL();
                    uint32_t e = lit[bb & 2047u];
                    if (e & kLiteralFlag) {
                        *op++ = (uint8_t)(e >> 16); FASTINF_DROP(e & 15u);
                        e = lit[bb & 2047u];
                        if (e & kLiteralFlag) {
                            *op++ = (uint8_t)(e >> 16); FASTINF_DROP(e & 15u);
                            e = lit[bb & 2047u];
                            if (e & kLiteralFlag) { *op++ = (uint8_t)(e >> 16); FASTINF_DROP(e & 15u); continue; }
                        }
                    }
                    if (((e >> kKindShift) & 7u) == 3u) { FASTINF_DROP(kLitBits); e = lit[(e >> 16) + (uint32_t)(bb & ((1u << ((e >> 8) & 15u)) - 1u))]; }
                    FASTINF_DROP(e & 15u);
                    const uint32_t kind = (e >> kKindShift) & 7u;
                    if (kind == 0) { *op++ = (uint8_t)(e >> 16); continue; }
                    if (kind == 2) { block_done = true; break; }
                    if (kind != 1) return -1;
                    FASTINF_REFILL();
                    unsigned xb = (e >> 8) & 15u;
                    const uint32_t length = (e >> 16) + (uint32_t)(bb & ((1u << xb) - 1u));
                    FASTINF_DROP(xb);
                    uint32_t d = dtab[bb & 255u];
                    if (((d >> kKindShift) & 7u) == 3u) { FASTINF_DROP(kDistBits); d = dtab[(d >> 16) + (uint32_t)(bb & ((1u << ((d >> 8) & 15u)) - 1u))]; }
                    if (((d >> kKindShift) & 7u) != 0) return -1;
                    FASTINF_DROP(d & 15u);
                    xb = (d >> 8) & 15u;
                    const uint32_t dist = (d >> 16) + (uint32_t)(bb & ((1u << xb) - 1u));
                    FASTINF_DROP(xb);
                    if (dist > (size_t)(op - window0)) return -1;
                    uint8_t* dst = op;
                    const uint8_t* src = op - dist;
                    op += length;
                    if (dist >= 8) {
                        uint64_t w;
                        memcpy(&w, src, 8); memcpy(dst, &w, 8);
                        memcpy(&w, src + 8, 8); memcpy(dst + 8, &w, 8);
                        if (length > 16) {
                            src += 16; dst += 16;
                            do { memcpy(&w, src, 8); memcpy(dst, &w, 8); src += 8; dst += 8; } while (dst < op);
                        }
                    } else if (dist == 1) {
                        memset(dst, *src, length);
                    } else {
                        for (uint32_t i = 0; i < length; ++i) dst[i] = src[i];
                    }
                }
#undef FASTINF_REFILL
#undef FASTINF_DROP
                r.buf = bb; r.n = bn; r.p = ip;
                o = (size_t)(op - out);
            }
            for (; !block_done;) {
                if (out_cap - o < 258 + 32 + 2 || r.past > 16) return -1;
                r.refill();
                uint32_t e = T->lit[r.peek(kLitBits)];
                if (e & kLiteralFlag) {                                  // up to three literals from one refill (a codeword has 15 bits at most)
                    out[o++] = (uint8_t)(e >> 16); r.drop(e & 15u);
                    e = T->lit[r.peek(kLitBits)];
                    if (e & kLiteralFlag) {
                        out[o++] = (uint8_t)(e >> 16); r.drop(e & 15u);
                        e = T->lit[r.peek(kLitBits)];
                        if (e & kLiteralFlag) { out[o++] = (uint8_t)(e >> 16); r.drop(e & 15u); continue; }
                    }
                }
                if (((e >> kKindShift) & 7u) == 3u) { r.drop(kLitBits); e = T->lit[(e >> 16) + r.peek((e >> 8) & 15u)]; }
                r.drop(e & 15u);
                const uint32_t kind = (e >> kKindShift) & 7u;
                if (kind == 0) { out[o++] = (uint8_t)(e >> 16); continue; }   // (a literal with a long codeword)
                if (kind == 2) break;
                if (kind != 1) return -1;
                r.refill();                                              // (unconditional: cheaper than a branch that depends on what came before)
                const uint32_t length = (e >> 16) + r.take((e >> 8) & 15u);
                uint32_t d = T->dist[r.peek(kDistBits)];
                if (((d >> kKindShift) & 7u) == 3u) { r.drop(kDistBits); d = T->dist[(d >> 16) + r.peek((d >> 8) & 15u)]; }
                if (((d >> kKindShift) & 7u) != 0) return -1;
                r.drop(d & 15u);
                const uint32_t dist = (d >> 16) + r.take((d >> 8) & 15u);
                if (dist > o - member_out) return -1;                // (a member's window starts with the member)
                uint8_t* dst = out + o;
                const uint8_t* src = dst - dist;
                o += length;
                if (dist >= 8) {                                        // sixteen bytes without asking (a match is 9 bytes on average), more in a loop
                    uint64_t w;
                    memcpy(&w, src, 8); memcpy(dst, &w, 8);
                    memcpy(&w, src + 8, 8); memcpy(dst + 8, &w, 8);
                    if (length > 16) {
                        uint8_t* const stop = dst + length;
                        src += 16; dst += 16;
                        do { memcpy(&w, src, 8); memcpy(dst, &w, 8); src += 8; dst += 8; } while (dst < stop);
                    }
                } else if (dist == 1) {
                    memset(dst, *src, length);
                } else {
                    for (uint32_t i = 0; i < length; ++i) dst[i] = src[i];
                }
            }
        } else return -1;
        if (r.past > 16) return -1;
        progress(o);
        if (!follow && o - member_out > ((size_t)4 << 20)) {
            try { follow = new FollowCrc(out, member_out); } catch (...) { follow = nullptr; }
        }
        if (follow) follow->target.store(o, std::memory_order_release);
        if (bfinal) return 1;
    }
}

// One gzip archive (all its members) from [in, in+in_len) to out (room for out_cap bytes, of which 32 may be scribbled on past the
// data).  `progress(bytes)` is called as the output grows (after every block).  Returns the output size, or (size_t)-1 when
// the decoder gave up -- *good then says how much output belongs to members that were completed and CRC-checked.
template <class Progress>
inline size_t gunzip(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_cap, size_t* good, Progress progress) {
    const size_t kFail = (size_t)-1;
    Tables* T = new Tables;
    struct Free { Tables* t; ~Free() { delete t; } } free_tables{T};
    size_t pos = 0, o = 0;
    *good = 0;
    bool first = true;
    while (pos < in_len) {
        // ---- member header (RFC 1952)
        if (in_len - pos < 18 || in[pos] != 0x1f || in[pos + 1] != 0x8b) { if (first) return kFail; break; }   // trailing garbage after a member: ignored, as zlib does
        if (in[pos + 2] != 8) return kFail;
        const uint8_t flg = in[pos + 3];
        if (flg & 0xE0) return kFail;
        size_t h = pos + 10;
        if (flg & 4) { if (h + 2 > in_len) return kFail; const size_t xlen = in[h] | ((size_t)in[h + 1] << 8); h += 2 + xlen; }
        if (flg & 8) { while (h < in_len && in[h]) ++h; ++h; }
        if (flg & 16) { while (h < in_len && in[h]) ++h; ++h; }
        if (flg & 2) h += 2;
        if (h >= in_len) return kFail;
        first = false;
        const size_t member_out = o;
        FollowCrc* follow = nullptr;
        struct DropFollow { FollowCrc*& f; ~DropFollow() { delete f; f = nullptr; } } drop_follow{follow};
        Reader r;
        r.p = in + h; r.end = in + in_len;
        // ---- deflate blocks (RFC 1951)
        if (inflate_blocks(r, in, T, out, o, out_cap, member_out, (size_t)-1, progress, follow) != 1) return kFail;
        // ---- member trailer: to a byte boundary, CRC-32 and length of the member's output
        r.drop(r.n & 7);
        const unsigned back = r.n >> 3;
        if (r.past > back) return kFail;
        const uint8_t* t = r.p - (back - r.past);
        if (r.end - t < 8) return kFail;
        const uint32_t crc = t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
        const uint32_t isize = t[4] | ((uint32_t)t[5] << 8) | ((uint32_t)t[6] << 16) | ((uint32_t)t[7] << 24);
        const size_t mlen = o - member_out;
        if ((uint32_t)mlen != isize) return kFail;
        uint32_t c = (uint32_t)crc32(0L, Z_NULL, 0);
        if (follow) c = follow->finish(o);
        else for (size_t at = 0; at < mlen;) { const size_t step = mlen - at < ((size_t)1 << 30) ? mlen - at : ((size_t)1 << 30); c = (uint32_t)crc32(c, out + member_out + at, (uInt)step); at += step; }
        if (c != crc) return kFail;
        *good = o;
        pos = (size_t)(t + 8 - in);
    }
    return o;
}

}  // namespace fastinf
