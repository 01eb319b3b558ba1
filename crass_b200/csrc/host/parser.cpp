// parser.cpp -- feed path: FASTA/FASTQ(.gz) -> byte-packed read batch.
//
// Reproduces the record stream that the reference's kseq_read() (src/crass/kseq.cpp:171-225)
// hands to searchFile (libcrispr.cpp:96-131), including its quirks:
//   * a record starts at the next '>' or '@'; the name ends at the first whitespace, the rest of
//     the line is the comment; mixed FASTA/FASTQ in one file is fine;
//   * the sequence is every isgraph() byte up to the next '>', '+' or '@' ANYWHERE (not only at a
//     line start); no case folding, no validation;
//   * a FASTQ quality string is read by LENGTH (so '@' inside it is harmless), one byte past the
//     last quality character is consumed, a length mismatch ends the stream with status -2;
//   * kseq never clears its comment/quality buffers: a record without a comment (or a FASTA record
//     after a FASTQ one) shows the previous record's string to searchFile ("stale" fields).  The
//     batch stores, per record, the offset of the string the reference would have seen.
//   * bytes are fetched as signed chars: 0xFF reads as EOF (-1) for the current loop only.
#include <ctype.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>

#include <stdexcept>

#include "internal.h"

namespace cbh {

Batch::~Batch() { if (bases) free_host(bases, pinned); }

void Batch::reserve_bases(size_t need) {
    if (need <= bases_cap) return;
    size_t ncap = bases_cap ? bases_cap : (1u << 20);
    while (ncap < need) ncap *= 2;
    bool pin = false;
    uint8_t* nb = (uint8_t*)alloc_host(ncap, &pin);
    if (!nb) throw std::bad_alloc();
    if (bases) {
        memcpy(nb, bases, (size_t)offsets.back());
        free_host(bases, pinned);
    }
    bases = nb; bases_cap = ncap; pinned = pin;
}

namespace {

bool inflate_all(const char* path, std::vector<uint8_t>& buf) {
    gzFile fp = (strcmp(path, "-") == 0) ? gzdopen(fileno(stdin), "r") : gzopen(path, "r");
    if (!fp) return false;
    gzbuffer(fp, 1 << 20);
    size_t n = 0;
    buf.resize(1 << 22);
    for (;;) {
        if (buf.size() - n < (1u << 21)) buf.resize(buf.size() * 2);
        size_t want = buf.size() - n;
        if (want > (1u << 30)) want = 1u << 30;
        int r = gzread(fp, buf.data() + n, (unsigned)want);
        if (r <= 0) break;
        n += (size_t)r;
    }
    gzclose(fp);
    buf.resize(n);
    return true;
}

struct Cursor {
    const uint8_t* p; size_t n, pos;
    int getc() { return pos < n ? (int)(signed char)p[pos++] : -1; }      // ks_getc (kseq.cpp:55-69)
};

inline bool c_isspace(uint8_t c) { return c == ' ' || (c >= '\t' && c <= '\r'); }
inline bool c_isgraph(int c) { return c > 32 && c < 127; }

// ks_getuntil (kseq.cpp:71-147); delimiter 0 == any whitespace.  Returns false when already at EOF
// (the reference returns -1 and leaves the target string untouched).
bool get_until(Cursor& c, int delimiter, size_t& b, size_t& e, int& dret) {
    dret = 0;
    if (c.pos >= c.n) return false;
    size_t i = c.pos;
    if (delimiter == 0) { while (i < c.n && !c_isspace(c.p[i])) ++i; }
    else {
        const void* q = memchr(c.p + i, delimiter, c.n - i);
        i = q ? (size_t)((const uint8_t*)q - c.p) : c.n;
    }
    b = c.pos; e = i;
    if (i < c.n) { dret = (int)(signed char)c.p[i]; c.pos = i + 1; } else c.pos = c.n;
    return true;
}

}  // namespace

int parse_file(const char* path, Batch** out) {
    std::vector<uint8_t> buf;
    if (!inflate_all(path, buf)) return fail(CRASS_B200_EIO, std::string("cannot open ") + path);
    Batch* B = new Batch();
    try {
        B->offsets.push_back(0);
        B->reserve_bases(buf.size() + 16);
        Cursor c{buf.data(), buf.size(), 0};
        int last_char = 0;
        int64_t cur_comment = -1, cur_qual = -1;
        uint64_t nb = 0;
        int status = -1;
        for (;;) {
            int ch;
            if (last_char == 0) {
                while ((ch = c.getc()) != -1 && ch != '>' && ch != '@') {}
                if (ch == -1) { status = -1; break; }
                last_char = ch;
            }
            size_t b, e; int dret;
            if (!get_until(c, 0, b, e, dret)) { status = -1; break; }
            const size_t name_b = b, name_e = e;
            if (dret != '\n') {
                size_t cb, ce; int d2;
                if (get_until(c, '\n', cb, ce, d2)) {
                    cur_comment = (int64_t)B->text_pool.size();
                    B->text_pool.insert(B->text_pool.end(), (const char*)buf.data() + cb, (const char*)buf.data() + ce);
                    B->text_pool.push_back(0);
                }
            }
            const uint64_t seq_b = nb;
            uint8_t* dst = B->bases;
            while ((ch = c.getc()) != -1 && ch != '>' && ch != '+' && ch != '@') {
                if (c_isgraph(ch)) dst[nb++] = (uint8_t)ch;
            }
            if (ch == '>' || ch == '@') last_char = ch;
            const uint64_t L = nb - seq_b;
            bool emit = true;
            if (ch == '+') {
                while ((ch = c.getc()) != -1 && ch != '\n') {}
                if (ch == -1) { status = -2; emit = false; }
                else {
                    const int64_t q0 = (int64_t)B->text_pool.size();
                    uint64_t ql = 0;
                    while ((ch = c.getc()) != -1 && ql < L) {
                        if (ch >= 33 && ch <= 127) { B->text_pool.push_back((char)ch); ++ql; }
                    }
                    B->text_pool.push_back(0);
                    cur_qual = q0;
                    last_char = 0;
                    if (ql != L) { status = -2; emit = false; }
                }
            }
            if (!emit) { nb = seq_b; break; }
            B->name_off.push_back(B->name_pool.size());
            B->name_pool.insert(B->name_pool.end(), (const char*)buf.data() + name_b, (const char*)buf.data() + name_e);
            B->name_pool.push_back(0);
            B->comment_off.push_back(cur_comment);
            B->qual_off.push_back(cur_qual);
            B->offsets.push_back(nb);
            if (L > B->max_len) B->max_len = (uint32_t)L;
        }
        B->parse_status = status;
    } catch (std::exception& ex) {
        delete B;
        return fail(CRASS_B200_ENOMEM, std::string("parse_file: ") + ex.what());
    }
    *out = B;
    return 0;
}

}  // namespace cbh
