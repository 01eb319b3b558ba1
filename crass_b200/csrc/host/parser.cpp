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
#include <fcntl.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
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

// The file's bytes: a read-only mapping for plain files (no copy at all), the inflated stream otherwise (gzip members,
// stdin).  zlib's transparent mode would read plain files too, but at a third of the speed of the parser behind it.
struct Input {
    std::vector<uint8_t> inflated;
    const uint8_t* data = nullptr;
    size_t size = 0;
    void* map = nullptr;
    size_t map_len = 0;
    ~Input() { if (map) munmap(map, map_len); }
    bool open(const char* path) {
        if (strcmp(path, "-") != 0) {
            const int fd = ::open(path, O_RDONLY);
            if (fd < 0) return false;
            struct stat st;
            unsigned char magic[2] = {0, 0};
            const bool regular = fstat(fd, &st) == 0 && S_ISREG(st.st_mode);
            const ssize_t got = regular ? pread(fd, magic, 2, 0) : 0;
            const bool gz = got == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
            if (regular && !gz) {
                if (st.st_size == 0) { ::close(fd); data = (const uint8_t*)""; size = 0; return true; }
                void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
                ::close(fd);
                if (m != MAP_FAILED) {
                    madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
                    map = m; map_len = (size_t)st.st_size; data = (const uint8_t*)m; size = map_len;
                    return true;
                }
            } else ::close(fd);
        }
        if (!inflate_all(path, inflated)) return false;
        data = inflated.data(); size = inflated.size();
        return true;
    }
};

struct Cursor {
    const uint8_t* p; size_t n, pos;
    int getc() { return pos < n ? (int)(signed char)p[pos++] : -1; }      // ks_getc (kseq.cpp:55-69)
};

inline bool c_isspace(uint8_t c) { return c == ' ' || (c >= '\t' && c <= '\r'); }
inline bool c_isgraph(int c) { return c > 32 && c < 127; }

// bytes a sequence loop appends without a second look: isgraph() and none of '>', '+', '@'
struct PlainTab {
    bool t[256];
    PlainTab() { for (int i = 0; i < 256; ++i) t[i] = c_isgraph(i) && i != '>' && i != '+' && i != '@'; }
};
const PlainTab kPlain;

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
    Input in;
    if (!in.open(path)) return fail(CRASS_B200_EIO, std::string("cannot open ") + path);
    Batch* B = new Batch();
    try {
        B->offsets.push_back(0);
        B->reserve_bases(in.size + 16);
        Cursor c{in.data, in.size, 0};
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
                    B->text_pool.insert(B->text_pool.end(), (const char*)in.data + cb, (const char*)in.data + ce);
                    B->text_pool.push_back(0);
                }
            }
            const uint64_t seq_b = nb;
            uint8_t* dst = B->bases;
            // kseq: while ((c = getc()) != -1 && c != '>' && c != '+' && c != '@') if (isgraph(c)) append(c);
            // taken in runs: bytes that are isgraph and none of the three terminators are copied in bulk, every other
            // byte is looked at on its own (0xFF reads as -1 and ends the loop like the others)
            for (;;) {
                size_t i = c.pos;
                while (i < c.n && kPlain.t[c.p[i]]) ++i;
                if (i > c.pos) { memcpy(dst + nb, c.p + c.pos, i - c.pos); nb += i - c.pos; c.pos = i; }
                ch = c.getc();
                if (ch == -1 || ch == '>' || ch == '+' || ch == '@') break;
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
                    // kseq: while ((c = getc()) != -1 && qual.l < seq.l) if (c >= 33 && c <= 127) append(c);  -- the byte
                    // is fetched before the length test, so one byte past the last quality character is consumed
                    while ((ch = c.getc()) != -1 && ql < L) {
                        if (ch < 33) continue;                             // (signed) also skips bytes >= 0x80
                        size_t i = c.pos;                                  // the rest of this run of quality characters
                        const size_t lim = c.pos + (size_t)(L - ql - 1) < c.n ? c.pos + (size_t)(L - ql - 1) : c.n;
                        while (i < lim && c.p[i] >= 33 && c.p[i] <= 127) ++i;
                        B->text_pool.push_back((char)ch);
                        B->text_pool.insert(B->text_pool.end(), (const char*)c.p + c.pos, (const char*)c.p + i);
                        ql += 1 + (i - c.pos);
                        c.pos = i;
                    }
                    B->text_pool.push_back(0);
                    cur_qual = q0;
                    last_char = 0;
                    if (ql != L) { status = -2; emit = false; }
                }
            }
            if (!emit) { nb = seq_b; break; }
            B->name_off.push_back(B->name_pool.size());
            B->name_pool.insert(B->name_pool.end(), (const char*)in.data + name_b, (const char*)in.data + name_e);
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
