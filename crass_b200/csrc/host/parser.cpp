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
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <atomic>
#include <chrono>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <thread>

#include "fast_inflate.h"
#include "internal.h"

namespace cbh {

Batch::~Batch() { if (bases) free_host(bases, pinned, registered); if (compact) free(compact); }

const uint8_t* Batch::contiguous() {
    if (segs.empty()) return bases;
    if (!compact) {
        const size_t total = offsets.empty() ? 0 : (size_t)offsets.back();
        compact = (uint8_t*)malloc(total + 16);
        if (!compact) throw std::bad_alloc();
        for (size_t k = 0; k < segs.size(); ++k) {
            const uint64_t end = k + 1 < segs.size() ? segs[k + 1].dev0 : (uint64_t)total;
            memcpy(compact + segs[k].dev0, bases + segs[k].host0, (size_t)(end - segs[k].dev0));
        }
    }
    return compact;
}

void Batch::reserve_bases(size_t need) {
    if (need <= bases_cap) return;
    size_t ncap = bases_cap ? bases_cap : (1u << 20);
    while (ncap < need) ncap *= 2;
    // large buffers in steps of 32 MB instead of powers of two: a 128 MB range needs 134 MB, and what an engine page-locks
    // (time and locked memory) should be what it uses
    if (need > ((size_t)64 << 20)) ncap = (need + ((size_t)32 << 20) - 1) & ~(((size_t)32 << 20) - 1);
    bool pin = false;
    uint8_t* nb = (uint8_t*)alloc_host(ncap, &pin, want_pinned);
    if (!nb) throw std::bad_alloc();
    if (bases) {
        if (!offsets.empty() && segs.empty()) memcpy(nb, bases, (size_t)offsets.back());
        free_host(bases, pinned, registered);
    }
    bases = nb; bases_cap = ncap; pinned = pin; registered = false;
}

namespace {

// zlib drops what a FAILING gzread call had inflated so far.  The reference reads 4096 bytes per call (kseq.cpp:44,60-66) and so
// keeps everything up to the 4096-byte chunk a damaged stretch of the archive falls into; a reader that asks for megabytes per
// call goes back over the failed call the reference's way.  `have` (a multiple of 4096) bytes are good; returns the new count.
size_t reread_like_kseq(const char* path, uint8_t* dst, size_t have, size_t cap, bool* ended_by_error) {
    *ended_by_error = false;
    gzFile fp = gzopen(path, "r");
    if (!fp) return have;
    if (gzseek(fp, (z_off_t)have, SEEK_SET) == (z_off_t)have) {
        while (have + 4096 <= cap) {
            // kstream reads into ONE buffer: when a call fails, what the buffer's first byte holds -- zlib may have copied part of
            // the chunk before it met the damage, else it is still the previous chunk's -- is what ks_getc hands out (see Cursor).
            // dst[have] plays that byte.
            if (have >= 4096) dst[have] = dst[have - 4096];
            const int r = gzread(fp, dst + have, 4096);
            if (r < 0) { *ended_by_error = true; break; }
            have += (size_t)r;
            if (r < 4096) break;
        }
    }
    gzclose(fp);
    return have;
}

bool inflate_all(const char* path, std::vector<uint8_t>& buf, bool* read_error) {
    *read_error = false;
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
        if (r < 0 && strcmp(path, "-") != 0 && n % 4096 == 0) { gzclose(fp); fp = nullptr; n = reread_like_kseq(path, buf.data(), n, buf.size(), read_error); break; }
        if (r <= 0) break;
        n += (size_t)r;
    }
    if (fp) gzclose(fp);
    const uint8_t after = n < buf.size() ? buf[n] : 0;
    buf.resize(n + 1);                                                    // (one byte past the input: what Cursor hands out after a failed read)
    buf[n] = after;
    return true;
}

size_t env_size_early(const char* name, size_t dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    char* end = nullptr;
    const unsigned long long x = strtoull(v, &end, 10);
    return end && *end == 0 ? (size_t)x : dflt;
}

// The file's bytes: a read-only mapping for plain files (no copy at all), the inflated stream otherwise (gzip members,
// stdin).  zlib's transparent mode would read plain files too, but at a third of the speed of the parser behind it.
struct Input {
    // gz inputs of a ParseStream are inflated by a thread of their own into a large reserved mapping while the ranges already
    // inflated are parsed, copied and searched (the reference inflates inside its read loop, SeqUtils.cpp:100-125 + kseq.cpp:55-69)
    std::thread inflater;
    std::atomic<size_t> avail{0};
    std::atomic<bool> inflate_done{false}, inflate_failed{false};
    std::atomic<bool> read_error{false};           // the input ended with a failed read (a damaged archive): see Cursor
    bool incremental = false;
    size_t vcap = 0;
    size_t expect_size = 0;                        // BGZF: the inflated size, known from the block trailers before anything is inflated
    // A BGZF archive (bgzip, htslib: gzip members of at most 64 KB whose extra field says how long each member is) can be inflated
    // by several threads: the block list is read off the headers, every block's place in the output is the sum of the ISIZE
    // trailers before it.  Any other gzip file is one deflate stream, and one thread (zlib) is all that can work on it.
    struct BgzfBlock { size_t cpos, clen, out; uint32_t isize, crc; size_t mpos, msize; };   // deflate payload, place in the output, the whole member
    void* cmap = nullptr; size_t cmap_len = 0;     // the archive itself, mapped (BGZF only)
    std::string bgzf_path;                         // set by open_streaming: a damaged block is then read again through zlib
    static bool bgzf_index(const uint8_t* c, size_t n, std::vector<BgzfBlock>& blocks, size_t* total) {
        size_t pos = 0, out = 0;
        while (pos < n) {
            if (n - pos < 28 || c[pos] != 0x1f || c[pos + 1] != 0x8b || c[pos + 2] != 8 || !(c[pos + 3] & 4)) return false;
            if (c[pos + 3] & ~4u) return false;                                  // a name, comment or header CRC: not what bgzip writes
            const size_t xlen = c[pos + 10] | ((size_t)c[pos + 11] << 8);
            if (pos + 12 + xlen > n) return false;
            size_t bsize = 0;
            for (size_t x = pos + 12; x + 4 <= pos + 12 + xlen;) {               // subfields: SI1 SI2 SLEN data
                const size_t slen = c[x + 2] | ((size_t)c[x + 3] << 8);
                if (c[x] == 'B' && c[x + 1] == 'C' && slen == 2 && x + 6 <= pos + 12 + xlen) bsize = (size_t)(c[x + 4] | (c[x + 5] << 8)) + 1;
                x += 4 + slen;
            }
            if (!bsize || bsize < 12 + xlen + 8 || pos + bsize > n) return false;
            const uint8_t* t = c + pos + bsize - 8;
            const uint32_t crc = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24);
            const uint32_t isize = t[4] | (t[5] << 8) | (t[6] << 16) | ((uint32_t)t[7] << 24);
            if (isize > (1u << 16)) return false;
            blocks.push_back(BgzfBlock{pos + 12 + xlen, bsize - 12 - xlen - 8, out, isize, crc, pos, bsize});
            out += isize;
            pos += bsize;
        }
        *total = out;
        return !blocks.empty();
    }
    // the inflater of a BGZF archive: worker threads take groups of blocks in file order; `avail` is the end of the leading run of
    // finished groups
    void inflate_bgzf(std::vector<BgzfBlock> blocks, unsigned n_threads) {
        const uint8_t* c = (const uint8_t*)cmap;
        const size_t group = 32, n_groups = (blocks.size() + group - 1) / group;
        std::vector<std::atomic<uint8_t> > done(n_groups);
        for (auto& d : done) d.store(0);
        std::atomic<size_t> next{0}, first_bad{(size_t)-1};
        std::mutex lead_mu;
        size_t lead = 0;
        auto work = [&]() {
            z_stream z;
            memset(&z, 0, sizeof z);
            if (inflateInit2(&z, -15) != Z_OK) { inflate_failed.store(true); return; }
            std::vector<uint8_t> tmp((size_t)65536 + 1024);
            const bool no_fast = getenv("CRASS_B200_GZ_ZLIB_BLOCKS") != nullptr;
            for (;;) {
                const size_t g = next.fetch_add(1);
                if (g >= n_groups || inflate_failed.load()) break;
                for (size_t b = g * group; b < std::min(blocks.size(), (g + 1) * group); ++b) {
                    const BgzfBlock& k = blocks[b];
                    uint8_t* dst = (uint8_t*)map + k.out;
                    // the own decoder first (it writes a little past its output, so it works in a buffer of the thread and the
                    // block is copied to its place); zlib for a block it does not accept
                    if (!no_fast) {
                        size_t good = 0;
                        const size_t got = fastinf::gunzip(c + k.mpos, k.msize, tmp.data(), tmp.size(), &good, [](size_t) {});
                        if (got == (size_t)k.isize) { memcpy(dst, tmp.data(), k.isize); continue; }
                    }
                    inflateReset(&z);
                    z.next_in = const_cast<Bytef*>(c + k.cpos); z.avail_in = (uInt)k.clen;
                    z.next_out = dst; z.avail_out = k.isize;
                    const int r = inflate(&z, Z_FINISH);
                    if (r != Z_STREAM_END || z.total_out != k.isize || (uint32_t)crc32(crc32(0L, Z_NULL, 0), dst, k.isize) != k.crc) {
                        size_t cur = first_bad.load();
                        while (b < cur && !first_bad.compare_exchange_weak(cur, b)) {}
                        inflate_failed.store(true);
                        break;
                    }
                }
                if (inflate_failed.load()) break;                                // (this group or another: nothing past a damaged block is handed out)
                done[g].store(1, std::memory_order_release);
                std::lock_guard<std::mutex> l(lead_mu);
                while (lead < n_groups && done[lead].load(std::memory_order_acquire)) ++lead;
                avail.store(lead < n_groups ? blocks[lead * group].out : expect_size, std::memory_order_release);
            }
            inflateEnd(&z);
        };
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < n_threads; ++t) pool.emplace_back(work);
        work();
        for (auto& t : pool) t.join();
        if (inflate_failed.load() && !bgzf_path.empty()) {
            // A damaged block.  Everything before it is in place; how far INTO it the reference would read depends on zlib's
            // buffering and kstream's 4096-byte calls, so that stretch is read again the reference's way (reread_like_kseq
            // inflates from the top of the archive to get there: slow, but only a damaged archive pays).
            const size_t bad = first_bad.load();
            size_t have = (bad < blocks.size() ? blocks[bad].out : expect_size) & ~(size_t)4095;
            bool by_error = false;
            have = reread_like_kseq(bgzf_path.c_str(), (uint8_t*)map, have, vcap, &by_error);
            read_error.store(by_error, std::memory_order_release);
            avail.store(have, std::memory_order_release);
            inflate_failed.store(false);
        }
        inflate_done.store(true, std::memory_order_release);
    }
    std::vector<uint8_t> inflated;
    const uint8_t* data = nullptr;
    size_t size = 0;
    void* map = nullptr;
    size_t map_len = 0;
    ~Input() {
        if (inflater.joinable()) inflater.join();
        if (map) unmap_later(map, map_len);
        if (cmap) munmap(cmap, cmap_len);
    }
    // open for streaming: plain files as open() does; gz files start an inflater thread and return at once
    bool open_streaming(const char* path) {
        if (strcmp(path, "-") == 0) return open(path);
        const int fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        unsigned char magic[2] = {0, 0};
        const bool regular = fstat(fd, &st) == 0 && S_ISREG(st.st_mode);
        const bool gz = regular && pread(fd, magic, 2, 0) == 2 && magic[0] == 0x1f && magic[1] == 0x8b;
        ::close(fd);
        if (!gz || (size_t)st.st_size < env_size_early("CRASS_B200_GZ_STREAM_MIN", (size_t)8 << 20)) return open(path);      // small archives: inflate first, as before
        if (!getenv("CRASS_B200_GZ_SERIAL")) {                                     // BGZF: every block on its own, several threads
            const int cfd = ::open(path, O_RDONLY);
            void* cm = cfd >= 0 ? mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, cfd, 0) : MAP_FAILED;
            if (cfd >= 0) ::close(cfd);
            if (cm != MAP_FAILED) {
                std::vector<BgzfBlock> blocks;
                size_t total = 0;
                if (bgzf_index((const uint8_t*)cm, (size_t)st.st_size, blocks, &total)) {
                    vcap = total + ((size_t)1 << 20);
                    void* m = mmap(nullptr, vcap, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
                    if (m != MAP_FAILED) {
                        madvise(m, vcap, MADV_HUGEPAGE);
                        cmap = cm; cmap_len = (size_t)st.st_size;
                        map = m; map_len = vcap; data = (const uint8_t*)m; size = 0; incremental = true; expect_size = total;
                        bgzf_path = path;
                        unsigned hw = std::thread::hardware_concurrency();
                        const unsigned nt = (unsigned)std::max<size_t>(1, env_size_early("CRASS_B200_GZ_THREADS", std::max<unsigned>(2, std::min<unsigned>(16, hw > 2 ? hw - 2 : 1))));
                        inflater = std::thread([this, nt](std::vector<BgzfBlock> bl) { inflate_bgzf(std::move(bl), nt); }, std::move(blocks));
                        return true;
                    }
                }
                munmap(cm, (size_t)st.st_size);
            }
        }
        vcap = (size_t)st.st_size * 64 + ((size_t)1 << 30);                        // address space only (MAP_NORESERVE)
        void* m = mmap(nullptr, vcap, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (m == MAP_FAILED) return open(path);
        madvise(m, vcap, MADV_HUGEPAGE);                                           // (a fault per 4 KB of output is a fifth of the inflater's time)
        map = m; map_len = vcap; data = (const uint8_t*)m; size = 0; incremental = true;
        const std::string p = path;
        // (tests shrink the inflater's step and the parser's margin so that ranges are cut while the stream is still arriving)
        const size_t step = std::min<size_t>((size_t)8 << 20, std::max<size_t>(4096, env_size_early("CRASS_B200_GZ_STREAM_MARGIN", (size_t)8 << 20))) & ~(size_t)4095;
        // An ordinary .gz is one deflate stream: one thread inflates it, and that thread is what the run waits for.  fast_inflate.h
        // does it at two to three times zlib's rate (no window, one table look-up per symbol, straight into the mapping); whatever it
        // does not accept -- above all a damaged or truncated archive -- is read again through zlib, which then decides what the
        // archive yields.  What is handed out while the decoder runs stays 32 KB behind it: less than zlib itself would have
        // delivered (its 16 KB buffer plus kstream's 4096-byte call) should the next block turn out to be damaged.
        void* cm = MAP_FAILED;
        if (!getenv("CRASS_B200_GZ_SERIAL")) {
            const int cfd = ::open(path, O_RDONLY);
            if (cfd >= 0) { cm = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, cfd, 0); ::close(cfd); }
        }
        if (cm != MAP_FAILED) {
            cmap = cm; cmap_len = (size_t)st.st_size;
            madvise(cm, cmap_len, MADV_SEQUENTIAL);
            inflater = std::thread([this, p]() {
                size_t good = 0;
                const size_t total = fastinf::gunzip((const uint8_t*)cmap, cmap_len, (uint8_t*)map, vcap, &good, [this](size_t o) {
                    if (o > (size_t)32768 && o - 32768 > avail.load(std::memory_order_relaxed)) avail.store(o - 32768, std::memory_order_release);
                });
                if (total != (size_t)-1) avail.store(total, std::memory_order_release);
                else {
                    bool by_error = false;
                    const size_t nread = reread_like_kseq(p.c_str(), (uint8_t*)map, good & ~(size_t)4095, vcap, &by_error);
                    if (nread + 4096 > vcap) inflate_failed.store(true);                 // more than 64 x the archive: give up loudly
                    read_error.store(by_error, std::memory_order_release);
                    if (nread > avail.load(std::memory_order_relaxed)) avail.store(nread, std::memory_order_release);
                    else if (nread < avail.load(std::memory_order_relaxed)) inflate_failed.store(true);   // (cannot happen, see above; never hand out less than was handed out)
                }
                inflate_done.store(true, std::memory_order_release);
            });
            return true;
        }
        inflater = std::thread([this, p, step]() {
            gzFile fp = gzopen(p.c_str(), "r");
            if (!fp) { inflate_failed.store(true); inflate_done.store(true); return; }
            gzbuffer(fp, 1 << 20);
            size_t nread = 0;
            for (;;) {
                const size_t want = std::min<size_t>(step, vcap - nread);
                if (!want) { inflate_failed.store(true); break; }                  // more than 64 x the archive: give up loudly
                const int r = gzread(fp, (uint8_t*)map + nread, (unsigned)want);
                if (r < 0 && nread % 4096 == 0) {                                  // a damaged archive: up to where the reference reads it
                    gzclose(fp); fp = nullptr;
                    bool by_error = false;
                    nread = reread_like_kseq(p.c_str(), (uint8_t*)map, nread, vcap, &by_error);
                    read_error.store(by_error, std::memory_order_release);
                    avail.store(nread, std::memory_order_release);
                    break;
                }
                if (r <= 0) break;
                nread += (size_t)r;
                avail.store(nread, std::memory_order_release);
            }
            if (fp) gzclose(fp);
            inflate_done.store(true, std::memory_order_release);
        });
        return true;
    }
    // blocks until `want` bytes are there or the stream has ended; returns what is there
    size_t wait_for(size_t want, bool* final_size) {
        if (!incremental) { *final_size = true; return size; }
        for (;;) {
            const bool d = inflate_done.load(std::memory_order_acquire);
            const size_t a = avail.load(std::memory_order_acquire);
            if (d) { size = a; *final_size = true; return a; }
            if (a >= want) { *final_size = false; return a; }
            std::this_thread::sleep_for(std::chrono::microseconds(200));
        }
    }
    // giving 1.6 GB of mapped pages back takes the kernel tens of milliseconds: nobody has to wait for it
    static void unmap_later(void* p, size_t len) {
        if (len < ((size_t)64 << 20)) { munmap(p, len); return; }
        try { std::thread([p, len] { munmap(p, len); }).detach(); } catch (...) { munmap(p, len); }
    }
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
        if (strcmp(path, "-") != 0 && !getenv("CRASS_B200_GZ_SERIAL") && (open_bgzf_now(path) || open_gz_now(path))) return true;
        bool by_error = false;
        if (!inflate_all(path, inflated, &by_error)) return false;
        read_error.store(by_error);
        data = inflated.data(); size = inflated.size() - 1;
        return true;
    }
    // an ordinary .gz, inflated by the own decoder before returning; whatever that does not accept is left to zlib (inflate_all)
    bool open_gz_now(const char* path) {
        const int cfd = ::open(path, O_RDONLY);
        struct stat st;
        if (cfd < 0 || fstat(cfd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size < 18) { if (cfd >= 0) ::close(cfd); return false; }
        void* cm = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, cfd, 0);
        ::close(cfd);
        if (cm == MAP_FAILED) return false;
        const size_t cap = (size_t)st.st_size * 64 + ((size_t)1 << 30);          // address space only
        void* m = mmap(nullptr, cap, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (m == MAP_FAILED) { munmap(cm, (size_t)st.st_size); return false; }
        if ((size_t)st.st_size >= ((size_t)1 << 20)) madvise(m, cap, MADV_HUGEPAGE);
        size_t good = 0;
        const size_t total = fastinf::gunzip((const uint8_t*)cm, (size_t)st.st_size, (uint8_t*)m, cap, &good, [](size_t) {});
        munmap(cm, (size_t)st.st_size);
        if (total == (size_t)-1) { munmap(m, cap); return false; }
        map = m; map_len = cap; data = (const uint8_t*)m; size = total;
        return true;
    }
    // a BGZF archive, inflated on several threads before returning (the whole-file form of what open_streaming starts)
    bool open_bgzf_now(const char* path) {
        const int cfd = ::open(path, O_RDONLY);
        struct stat st;
        if (cfd < 0 || fstat(cfd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size < 28) { if (cfd >= 0) ::close(cfd); return false; }
        void* cm = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, cfd, 0);
        ::close(cfd);
        if (cm == MAP_FAILED) return false;
        std::vector<BgzfBlock> blocks;
        size_t total = 0;
        if (!bgzf_index((const uint8_t*)cm, (size_t)st.st_size, blocks, &total) || total < ((size_t)1 << 20)) { munmap(cm, (size_t)st.st_size); return false; }
        const size_t len = total + 4096;
        void* m = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (m == MAP_FAILED) { munmap(cm, (size_t)st.st_size); return false; }
        cmap = cm; cmap_len = (size_t)st.st_size; map = m; map_len = len; expect_size = total;
        unsigned hw = std::thread::hardware_concurrency();
        inflate_bgzf(std::move(blocks), (unsigned)std::max<size_t>(1, env_size_early("CRASS_B200_GZ_THREADS", std::max<unsigned>(2, std::min<unsigned>(16, hw > 2 ? hw - 2 : 1)))));
        if (inflate_failed.load()) {                                             // a damaged block: let zlib say how far the archive reads
            munmap(map, map_len); map = nullptr; map_len = 0;
            munmap(cmap, cmap_len); cmap = nullptr; cmap_len = 0;
            inflate_failed.store(false); inflate_done.store(false); avail.store(0); expect_size = 0;
            return false;
        }
        data = (const uint8_t*)map; size = total;
        return true;
    }
};

// kstream refills a 4096-byte buffer and learns about the end of the input from a SHORT read: when the input's size is a
// multiple of 4096 that flag is still clear after the last byte has been consumed, and the first ks_getuntil called there
// returns an empty string instead of -1 (kseq.cpp:71-92) -- one extra record with an empty name and an empty sequence when
// the last byte of such an input is a header character.  With the whole input in memory that state is one bit.
// A FAILED read (gzread returns -1 on a damaged archive) is taken for a short one by kstream: end = -1 sets is_eof, but only
// end == 0 makes ks_getc return -1 -- the ks_getc that meets the failure hands out buf[0] of the last good 4096-byte chunk once
// more before the stream ends (or, if zlib had copied part of the failing chunk into the buffer, that chunk's first byte), while a
// ks_getuntil in that place copies nothing and leaves begin = 1 (kseq.cpp:55-69,84-96).
struct Cursor {
    const uint8_t* p; size_t n, pos;
    bool is_eof;
    bool read_error = false, stale_done = false;
    Cursor(const uint8_t* p_, size_t n_, size_t pos_) : p(p_), n(n_), pos(pos_), is_eof(n_ % 4096 != 0) {}
    int getc() {                                                          // ks_getc (kseq.cpp:55-69)
        if (pos < n) return (int)(signed char)p[pos++];
        if (read_error && !stale_done && !is_eof && n >= 4096) { stale_done = true; is_eof = true; return (int)(signed char)p[n]; }   // (reread_like_kseq left it there)
        is_eof = true;
        return -1;
    }
};

inline bool c_isspace(uint8_t c) { return c == ' ' || (c >= '\t' && c <= '\r'); }
inline bool c_isgraph(int c) { return c > 32 && c < 127; }

// bytes a sequence loop appends without a second look: isgraph() and none of '>', '+', '@'
struct PlainTab {
    bool t[256];
    PlainTab() { for (int i = 0; i < 256; ++i) t[i] = c_isgraph(i) && i != '>' && i != '+' && i != '@'; }
};
const PlainTab kPlain;

// bytes [b, e) plus a terminating NUL appended to a pool with one memcpy (vector::insert with a custom allocator constructs
// element by element)
inline void append_str(PodVec<char>& v, const char* b, const char* e, bool nul = true) {
    const size_t at = v.size(), len = (size_t)(e - b);
    if (v.capacity() < at + len + 1) v.reserve(std::max<size_t>(v.capacity() * 2, at + len + 1 + 4096));
    v.resize(at + len + (nul ? 1 : 0));
    memcpy(v.data() + at, b, len);
    if (nul) v[at + len] = 0;
}

// the end of the run of plain bytes that starts at i (first byte that is not plain, or lim), 16 bytes at a time:
// not plain == below 33 as a SIGNED byte (control, blank, and everything from 0x80 up), 127, or one of '>' '+' '@'
inline size_t plain_run_end(const uint8_t* p, size_t i, size_t lim) {
#if defined(__SSE2__)
    const __m128i k33 = _mm_set1_epi8(33), k127 = _mm_set1_epi8(127), kgt = _mm_set1_epi8('>'), kplus = _mm_set1_epi8('+'), kat = _mm_set1_epi8('@');
    while (i + 16 <= lim) {
        const __m128i x = _mm_loadu_si128((const __m128i*)(p + i));
        const __m128i bad = _mm_or_si128(_mm_or_si128(_mm_cmplt_epi8(x, k33), _mm_cmpeq_epi8(x, k127)),
                                         _mm_or_si128(_mm_cmpeq_epi8(x, kgt), _mm_or_si128(_mm_cmpeq_epi8(x, kplus), _mm_cmpeq_epi8(x, kat))));
        const int m = _mm_movemask_epi8(bad);
        if (m) return i + (size_t)__builtin_ctz((unsigned)m);
        i += 16;
    }
#endif
    while (i < lim && kPlain.t[p[i]]) ++i;
    return i;
}

// the end of the run of quality characters (33..127) that starts at i
inline size_t qual_run_end(const uint8_t* p, size_t i, size_t lim) {
#if defined(__SSE2__)
    const __m128i k33 = _mm_set1_epi8(33);
    while (i + 16 <= lim) {
        const int m = _mm_movemask_epi8(_mm_cmplt_epi8(_mm_loadu_si128((const __m128i*)(p + i)), k33));
        if (m) return i + (size_t)__builtin_ctz((unsigned)m);
        i += 16;
    }
#endif
    while (i < lim && p[i] >= 33 && p[i] <= 127) ++i;
    return i;
}

// ks_getuntil (kseq.cpp:71-147); delimiter 0 == any whitespace.  Returns false when already at EOF
// (the reference returns -1 and leaves the target string untouched).
bool get_until(Cursor& c, int delimiter, size_t& b, size_t& e, int& dret) {
    dret = 0;
    if (c.pos >= c.n) {
        if (c.is_eof) return false;
        c.is_eof = true;                                                  // the refill that finds nothing happens inside this call
        c.stale_done = true;
        b = e = c.n;
        return true;
    }
    size_t i = c.pos;
    if (delimiter == 0) {
#if defined(__SSE2__)
        // isspace: ' ' or '\t'..'\r', 16 bytes at a time (a name is rarely longer)
        const __m128i ksp = _mm_set1_epi8(' '), k8 = _mm_set1_epi8(8), k14 = _mm_set1_epi8(14);
        while (i + 16 <= c.n) {
            const __m128i x = _mm_loadu_si128((const __m128i*)(c.p + i));
            const int m = _mm_movemask_epi8(_mm_or_si128(_mm_cmpeq_epi8(x, ksp), _mm_and_si128(_mm_cmpgt_epi8(x, k8), _mm_cmplt_epi8(x, k14))));
            if (m) { i += (size_t)__builtin_ctz((unsigned)m); goto found_space; }
            i += 16;
        }
#endif
        while (i < c.n && !c_isspace(c.p[i])) ++i;
    }
    else {
        const void* q = memchr(c.p + i, delimiter, c.n - i);
        i = q ? (size_t)((const uint8_t*)q - c.p) : c.n;
    }
#if defined(__SSE2__)
found_space:
#endif
    b = c.pos; e = i;
    if (i < c.n) { dret = (int)(signed char)c.p[i]; c.pos = i + 1; } else { c.pos = c.n; c.is_eof = true; c.stale_done = true; }
    return true;
}

// One stretch of the file parsed on its own: records whose header character lies in [start, stop).  The strings are kept
// in piece-local pools and spliced into the batch afterwards.  A piece never looks at what came before `start`, so the
// only state it inherits is kseq's stale comment/quality (offset -1 below = "whatever the previous piece left").
struct Piece {
    uint8_t* bases = nullptr; size_t cap = 0; bool own = false;
    uint64_t nb = 0;
    PodVec<uint64_t> ends;                       // end of each record's bases, piece-relative
    PodVec<char> name_pool, text_pool;
    PodVec<uint64_t> name_off;
    PodVec<int64_t> comment_off, qual_off;       // piece-relative; -1 = inherited
    int64_t last_comment = -1, last_qual = -1;
    uint32_t max_len = 0;
    int status = 0;                              // 0: stopped at a header >= stop (next_hp); -1/-2: the stream ended here
    size_t next_hp = 0;
    size_t from = 0;                             // file position the piece starts at (where its bases lie in a shared buffer)
    // comments and qualities: the piece's own pool, or -- for a streamed range -- its slice of the batch's text pool (a piece's
    // strings are no longer than its bytes); a slice that fills up (the last record ran past the piece's end, or the input
    // ended without a newline) is continued in the own pool and copied into place by the splice
    char* text_ext = nullptr; size_t text_n = 0, text_cap = 0;
    bool text_diverted = false;
    size_t text_size() const { return text_ext ? text_n : text_pool.size(); }
    const char* text_data() const { return text_ext ? text_ext : text_pool.data(); }
    void text_append(const char* b, const char* e, bool nul) {
        const size_t len = (size_t)(e - b), add = len + (nul ? 1 : 0);
        if (text_ext) {
            if (text_n + add <= text_cap) {
                if (len) memcpy(text_ext + text_n, b, len);
                if (nul) text_ext[text_n + len] = 0;
                text_n += add;
                return;
            }
            text_pool.clear();
            append_str(text_pool, text_ext, text_ext + text_n, false);
            text_ext = nullptr; text_diverted = true;
        }
        append_str(text_pool, b, e, nul);
    }
    ~Piece() { if (own) free(bases); }
    void rewind() {                              // empty, with the memory kept (a piece of a ParseArena, range after range)
        if (own) free(bases);
        own = false; bases = nullptr; cap = 0;
        nb = 0; ends.clear(); name_pool.clear(); text_pool.clear(); name_off.clear(); comment_off.clear(); qual_off.clear();
        last_comment = last_qual = -1; max_len = 0; status = 0; next_hp = 0;
        text_ext = nullptr; text_n = text_cap = 0; text_diverted = false;
    }
    void release() {                             // give everything back now (called by the thread that has just spliced the piece)
        if (own) free(bases);
        bases = nullptr; cap = 0;
        PodVec<uint64_t>().swap(ends); PodVec<char>().swap(name_pool); PodVec<char>().swap(text_pool);
        PodVec<uint64_t>().swap(name_off); PodVec<int64_t>().swap(comment_off); PodVec<int64_t>().swap(qual_off);
    }
    // room for the records of `span` input bytes without growing (address space only: pages are touched as they are written);
    // shorter records than assumed here grow the vectors as usual
    void expect(size_t span) {
        const size_t recs = span / 48 + 16;
        ends.reserve(recs); name_off.reserve(recs); comment_off.reserve(recs); qual_off.reserve(recs);
        name_pool.reserve(span / 6 + 64);
        text_span = span;
    }
    size_t text_span = 0;
    // room for `need` bases, `have` of which are written.  A piece that writes into its slice of a shared buffer fills it only
    // when its last record runs past the piece's end (the next piece's start was guessed inside a record): it goes on in a
    // buffer of its own, which the splice copies to where the piece belongs.
    void room(size_t need, size_t have) {
        if (need <= cap) return;
        size_t ncap = cap * 2 > need ? cap * 2 : need;
        if (!own) {
            uint8_t* nbuf = (uint8_t*)malloc(ncap);
            if (!nbuf) throw std::bad_alloc();
            if (have) memcpy(nbuf, bases, have);
            bases = nbuf; cap = ncap; own = true;
            return;
        }
        uint8_t* nbuf = (uint8_t*)realloc(bases, ncap);
        if (!nbuf) throw std::bad_alloc();
        bases = nbuf; cap = ncap;
    }
};

// The reference's read loop (libcrispr.cpp:96 over kseq_read, kseq.cpp:171-225) from the header character at `start`
// (start == SIZE_MAX: from the top of the file, looking for the first header) until a header at or past `stop`.
void parse_span(const uint8_t* data, size_t n, size_t start, size_t stop, Piece& P, bool read_error = false) {
    Cursor c(data, n, 0);
    c.read_error = read_error;
    int last_char = 0;
    if (start != (size_t)-1) { c.pos = start + 1; last_char = (int)(signed char)data[start]; }
    int64_t cur_comment = -1, cur_qual = -1;
    uint64_t nb = 0;
    for (;;) {
        int ch;
        if (last_char == 0) {
            while ((ch = c.getc()) != -1 && ch != '>' && ch != '@') {}
            if (ch == -1) { P.status = -1; break; }
            last_char = ch;
        }
        if (c.pos - 1 >= stop) { P.status = 0; P.next_hp = c.pos - 1; break; }
        size_t b, e; int dret;
        if (!get_until(c, 0, b, e, dret)) { P.status = -1; break; }
        const size_t name_b = b, name_e = e;
        if (dret != '\n') {
            size_t cb, ce; int d2;
            if (get_until(c, '\n', cb, ce, d2)) {
                cur_comment = (int64_t)P.text_size();
                P.text_append((const char*)data + cb, (const char*)data + ce, true);
            }
        }
        const uint64_t seq_b = nb;
        // kseq: while ((c = getc()) != -1 && c != '>' && c != '+' && c != '@') if (isgraph(c)) append(c);
        // taken in runs: bytes that are isgraph and none of the three terminators are copied in bulk, every other
        // byte is looked at on its own (0xFF reads as -1 and ends the loop like the others)
        for (;;) {
            const size_t i = plain_run_end(c.p, c.pos, c.n);
            if (i > c.pos) {
                P.room(nb + (i - c.pos), nb);
                memcpy(P.bases + nb, c.p + c.pos, i - c.pos); nb += i - c.pos; c.pos = i;
            }
            const bool at_end = c.pos >= c.n;
            ch = c.getc();
            if (ch == -1 || ch == '>' || ch == '+' || ch == '@') break;
            if (at_end && c_isgraph(ch)) { P.room(nb + 1, nb); P.bases[nb++] = (uint8_t)ch; }   // (the byte a failed read hands out again, see Cursor)
        }
        if (ch == '>' || ch == '@') last_char = ch;
        const uint64_t L = nb - seq_b;
        bool emit = true;
        int bad = 0;
        if (ch == '+') {
            while ((ch = c.getc()) != -1 && ch != '\n') {}
            if (ch == -1) { bad = -2; emit = false; }
            else {
                if (!P.text_ext && P.text_span && P.text_pool.capacity() < P.text_span / 2) P.text_pool.reserve(P.text_span / 2 + P.text_span / 8 + 64);
                const int64_t q0 = (int64_t)P.text_size();
                uint64_t ql = 0;
                // kseq: while ((c = getc()) != -1 && qual.l < seq.l) if (c >= 33 && c <= 127) append(c);  -- the byte
                // is fetched before the length test, so one byte past the last quality character is consumed
                for (;;) {
                    const bool at_end = c.pos >= c.n;
                    if ((ch = c.getc()) == -1 || !(ql < L)) break;
                    if (ch < 33) continue;                             // (signed) also skips bytes >= 0x80
                    if (at_end) { const char one = (char)ch; P.text_append(&one, &one + 1, false); ql += 1; continue; }   // (see Cursor)
                    const size_t lim = c.pos + (size_t)(L - ql - 1) < c.n ? c.pos + (size_t)(L - ql - 1) : c.n;
                    const size_t i = qual_run_end(c.p, c.pos, lim);     // the rest of this run of quality characters
                    P.text_append((const char*)c.p + c.pos - 1, (const char*)c.p + i, false);   // (ch is the byte before pos)
                    ql += 1 + (i - c.pos);
                    c.pos = i;
                }
                P.text_append(nullptr, nullptr, true);
                cur_qual = q0;
                last_char = 0;
                if (ql != L) { bad = -2; emit = false; }
            }
        }
        if (!emit) { nb = seq_b; P.status = bad; break; }
        P.name_off.push_back(P.name_pool.size());
        append_str(P.name_pool, (const char*)data + name_b, (const char*)data + name_e);
        P.comment_off.push_back(cur_comment);
        P.qual_off.push_back(cur_qual);
        P.ends.push_back(nb);
        if (L > P.max_len) P.max_len = (uint32_t)L;
    }
    P.nb = nb; P.last_comment = cur_comment; P.last_qual = cur_qual;
}

// A likely record start at or after `from` (and before `lim`): a '>' or '@' at a line start whose next line is not a
// header itself (a quality line may begin with either character) and, for '@', whose third line begins with '+'.
// Only a guess -- parse_file() checks every guess against the piece before it.
size_t guess_record_start(const uint8_t* d, size_t n, size_t from, size_t lim) {
    // CRASS_B200_PARSE_GUESS=naive (tests): the next '>' or '@' byte, wherever it is -- wrong most of the time in FASTQ,
    // which is what exercises the check-and-re-parse side of parse_file()
    static const bool naive = getenv("CRASS_B200_PARSE_GUESS") && strcmp(getenv("CRASS_B200_PARSE_GUESS"), "naive") == 0;
    if (naive) {
        for (size_t j = from; j < lim; ++j) if (d[j] == '>' || d[j] == '@') return j;
        return (size_t)-1;
    }
    size_t i = from;
    while (i < lim) {
        const void* q = memchr(d + i, '\n', lim - i);
        if (!q) break;
        const size_t h = (size_t)((const uint8_t*)q - d) + 1;
        i = h;
        if (h >= lim || h >= n || (d[h] != '>' && d[h] != '@')) continue;
        const void* e1 = memchr(d + h, '\n', n - h);
        if (!e1) break;
        const size_t l2 = (size_t)((const uint8_t*)e1 - d) + 1;
        if (l2 >= n || d[l2] == '>' || d[l2] == '@') continue;
        if (d[h] == '>') return h;
        const void* e2 = memchr(d + l2, '\n', n - l2);
        if (!e2) break;
        const size_t l3 = (size_t)((const uint8_t*)e2 - d) + 1;
        if (l3 < n && d[l3] == '+') return h;
    }
    return (size_t)-1;
}

size_t env_size(const char* name, size_t dflt) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt;
    char* end = nullptr;
    const unsigned long long x = strtoull(v, &end, 10);
    return end && *end == 0 && x > 0 ? (size_t)x : dflt;
}

}  // namespace

// Large inputs are cut at guessed record starts and the pieces parsed by worker threads (CRASS_B200_PARSE_THREADS,
// default min(hardware threads, 16); pieces of CRASS_B200_PARSE_CHUNK bytes, default 16 MiB).  kseq's stream has no
// resynchronisation point that can be recognised locally ('@' and '>' are legal quality characters), so a piece is
// only kept when the piece before it ENDS exactly on the header it started from; otherwise the gap is parsed again
// from the true position on the calling thread.  The record stream is therefore the sequential one by construction.
//
// parse_range() is that for the records whose header lies in [start, about stop_hint): the unit of the streamed feed
// (ParseStream below).  start == SIZE_MAX: from the top of the file; stop_hint == SIZE_MAX: to the end.  A range ends on a
// TRUE record start (out->next_hp, where the next range begins); what kseq carries from record to record -- the stale
// comment and quality strings -- travels in the RangeCarry.
namespace {
struct RangeCarry {
    size_t next_hp = (size_t)-1;
    bool has_comment = false, has_qual = false;
    std::string comment, qual;
    int status = 0;                                       // 0: more records follow; -1 / -2: the stream ended in this range
};

struct View { const uint8_t* data; size_t size; bool read_error = false; };     // the input's bytes as far as a range may look (read_error: they end with a failed read)

// What a streamed input keeps from one range to the next: the scratch mapping the pieces write their bases to and the pieces'
// own vectors.  Fresh memory for them (a few hundred MB per range) costs a page fault per 4 KB on the way in and a munmap on the
// way out, which for FASTQ (half the bytes are quality strings that go through a pool) was half the parsing time.
struct ParseArena {
    uint8_t* scratch = nullptr; size_t scratch_len = 0;
    std::vector<std::unique_ptr<Piece> > pieces;
    ~ParseArena() { if (scratch) Input::unmap_later(scratch, scratch_len); }
    uint8_t* room(size_t need) {
        if (need <= scratch_len) return scratch;
        if (scratch) Input::unmap_later(scratch, scratch_len);
        scratch = nullptr; scratch_len = 0;
        const size_t len = need + need / 8;
        void* sm = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (sm == MAP_FAILED) throw std::bad_alloc();
        madvise(sm, len, MADV_HUGEPAGE);
        scratch = (uint8_t*)sm; scratch_len = len;
        return scratch;
    }
};

void parse_range(const View& in, size_t start, size_t stop_hint, const RangeCarry* cin, Batch* B, RangeCarry* cout, ParseArena* arena = nullptr) {
    B->reset();
    B->offsets.push_back(0);
    const size_t n = in.size;
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    const size_t n_threads = env_size("CRASS_B200_PARSE_THREADS", hw < 16 ? hw : 16);
    size_t chunk = env_size("CRASS_B200_PARSE_CHUNK", (size_t)16 << 20);
    const bool whole = start == (size_t)-1 && stop_hint == (size_t)-1 && !cin && !cout;
    const size_t first = start == (size_t)-1 ? 0 : start;
    // the range's end: a guessed record start at or after stop_hint (SIZE_MAX: the range runs to the end of the input)
    size_t range_stop = (size_t)-1;
    if (stop_hint != (size_t)-1 && stop_hint < n) range_stop = guess_record_start(in.data, n, stop_hint, n);
    const size_t limit = range_stop == (size_t)-1 ? n : range_stop;
    // a range is a few hundred MB: at least two pieces per thread, so that the threads finish together
    if (!whole && !getenv("CRASS_B200_PARSE_CHUNK")) chunk = std::max<size_t>((size_t)1 << 20, std::min(chunk, (limit - first) / (2 * n_threads) + 1));
    std::vector<size_t> starts;                           // header positions the pieces 1.. start from
    if (n_threads > 1 && limit - first > chunk) {
        for (size_t at = first + chunk; at < limit; at += chunk) {
            const size_t s = guess_record_start(in.data, n, at, at + chunk < limit ? at + chunk : limit);
            if (s != (size_t)-1 && s < limit && (starts.empty() || s > starts.back())) starts.push_back(s);
        }
    }
    if (starts.empty() && whole) {                        // one piece, written straight into the batch
        B->reserve_bases(n + 16);
        Piece P;
        P.bases = B->bases; P.cap = B->bases_cap;
        parse_span(in.data, n, (size_t)-1, (size_t)-1, P, in.read_error);
        B->name_pool.swap(P.name_pool); B->text_pool.swap(P.text_pool);   // piece-relative == batch-relative here
        B->name_off.swap(P.name_off); B->comment_off.swap(P.comment_off); B->qual_off.swap(P.qual_off);
        B->offsets.reserve(P.ends.size() + 1);
        B->offsets.insert(B->offsets.end(), P.ends.begin(), P.ends.end());
        B->max_len = P.max_len;
        B->parse_status = P.status;
        return;
    }
    const size_t np = starts.size() + 1;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    // one anonymous mapping for the bases of all pieces (huge pages when the kernel grants them: a 16 MB malloc per piece
    // was 400 k page faults on the way in and as many pages to give back), unmapped on a detached thread at the end -- or, for
    // a streamed input, the arena's, kept from range to range
    //
    // A streamed range (arena given) goes one step further: its pieces write straight into the batch's own (page-locked) base
    // buffer, each at its file offset -- a piece's bases are fewer than its bytes, so the slices cannot collide -- and STAY
    // there: the batch records the segments, the device copy puts them back to back (Batch::segs).  No splice of the bases.
    ParseArena local_arena;
    ParseArena& A = arena ? *arena : local_arena;
    const bool in_place = arena != nullptr;
    uint8_t* scratch;
    // (the strings the range before left behind open this batch's text pool)
    const size_t carried = (cin && cin->has_comment ? cin->comment.size() + 1 : 0) + (cin && cin->has_qual ? cin->qual.size() + 1 : 0);
    if (in_place) {
        B->reserve_bases((limit - first) + 64);
        scratch = B->bases;
        B->text_pool.resize(carried + (limit - first) + 64);           // (address space: pages are touched as they are written)
    } else scratch = A.room((limit - first) + 64 * np + 4096);
    char* const text0 = in_place ? B->text_pool.data() + carried : nullptr;
    while (A.pieces.size() < np) A.pieces.emplace_back(new Piece());
    for (size_t k = 0; k < np; ++k) A.pieces[k]->rewind();
    std::vector<std::unique_ptr<Piece> >& pieces = A.pieces;
    std::atomic<size_t> next{0};
    std::atomic<bool> failed{false};
    auto work = [&]() {
        for (;;) {
            const size_t k = next.fetch_add(1);
            if (k >= np || failed.load()) return;
            try {
                Piece& P = *pieces[k];
                const size_t from = k ? starts[k - 1] : first, stop = k + 1 < np ? starts[k] : range_stop;
                // a piece's bases are a subset of its bytes: its slice of the scratch mapping, at its own file offset
                P.own = false;
                P.from = from;
                P.cap = ((stop == (size_t)-1 ? n : stop) - from) + (in_place ? 0 : 64);
                P.bases = scratch + (from - first) + (in_place ? 0 : 64 * k);
                if (in_place) { P.text_ext = text0 + (from - first); P.text_cap = P.cap; }
                P.expect(P.cap);
                parse_span(in.data, n, k ? from : start, stop, P, in.read_error);
            } catch (...) { failed.store(true); }
        }
    };
    {
        std::vector<std::thread> pool;
        const size_t nt = n_threads < np ? n_threads : np;
        for (size_t t = 1; t < nt; ++t) pool.emplace_back(work);
        work();
        for (auto& t : pool) t.join();
    }
    if (failed.load()) throw std::bad_alloc();

    const double t1 = now();
    // walk the pieces in file order; `order` lists what the sequential reader would have produced
    std::vector<const Piece*> order;
    std::vector<std::unique_ptr<Piece>> patches;
    size_t k = 0;
    const Piece* cur = pieces[0].get();
    int status = 0;
    size_t end_hp = (size_t)-1;
    for (;;) {
        order.push_back(cur);
        if (cur->status != 0) { status = cur->status; break; }
        const size_t hp = cur->next_hp;
        if (hp >= range_stop) { end_hp = hp; break; }                 // the range is complete: the next one starts at hp
        while (k < starts.size() && starts[k] < hp) ++k;          // pieces that began inside a record
        if (k < starts.size() && starts[k] == hp) { cur = pieces[++k].get(); continue; }
        std::unique_ptr<Piece> Q(new Piece());                    // no piece begins here: parse up to the next one
        const size_t stop = k < starts.size() ? starts[k] : range_stop;
        Q->from = hp;
        if (in_place) {                                           // the slices of the pieces that began inside a record are free
            Q->own = false;
            Q->cap = (stop == (size_t)-1 ? n : stop) - hp;
            Q->bases = scratch + (hp - first);
            Q->text_ext = text0 + (hp - first); Q->text_cap = Q->cap;
        } else {
            Q->own = true;
            Q->cap = ((stop == (size_t)-1 ? n : stop) - hp) + 64;
            Q->bases = (uint8_t*)malloc(Q->cap);
            if (!Q->bases) throw std::bad_alloc();
        }
        parse_span(in.data, n, hp, stop, *Q, in.read_error);
        cur = Q.get();
        patches.push_back(std::move(Q));
    }
    // where each piece lands in the batch, and the stale comment/quality it inherits (a range inherits the strings the
    // range before it left: they open this batch's text pool)
    struct Slot { uint64_t base0, rec0, name0; int64_t text0, in_comment, in_qual; };
    std::vector<Slot> slot(order.size());
    uint64_t total = 0, n_rec = 0, n_name = 0;
    int64_t n_text = 0, cc = -1, cq = -1;
    if (cin && cin->has_comment) { cc = n_text; n_text += (int64_t)cin->comment.size() + 1; }
    if (cin && cin->has_qual) { cq = n_text; n_text += (int64_t)cin->qual.size() + 1; }
    const int64_t carried_text = n_text;
    for (size_t i = 0; i < order.size(); ++i) {
        const Piece& P = *order[i];
        // (in place, a piece's strings are -- or, if it had to go on in its own pool, will be -- in its slice of the text pool)
        const int64_t t0p = in_place ? (int64_t)(carried + (P.from - first)) : n_text;
        slot[i] = Slot{total, n_rec, n_name, t0p, cc, cq};
        if (P.last_comment >= 0) cc = t0p + P.last_comment;
        if (P.last_qual >= 0) cq = t0p + P.last_qual;
        total += P.nb; n_rec += P.ends.size(); n_name += P.name_pool.size();
        n_text = in_place ? std::max<int64_t>(n_text, t0p + (int64_t)P.text_size()) : n_text + (int64_t)P.text_size();
        if (P.max_len > B->max_len) B->max_len = P.max_len;
    }
    if (in_place) n_text = std::max<int64_t>(n_text, (int64_t)B->text_pool.size());   // (never shrunk: slices of later pieces lie further up)
    const double t2 = now();
    // in place: a piece that went on in a buffer of its own (room()) is copied to where it belongs, which is free by now (what it
    // overran were slices of pieces that began inside its last record); should the last record of the range run past the
    // buffer, everything moves into a larger one, back to back
    if (in_place) {
        size_t need_end = 0;
        for (const Piece* P : order) need_end = std::max(need_end, (P->from - first) + (size_t)P->nb);
        if (need_end + 16 > B->bases_cap) {
            bool pin = false;
            uint8_t* nbuf = (uint8_t*)alloc_host(total + 64, &pin, B->want_pinned);
            if (!nbuf) throw std::bad_alloc();
            for (size_t i = 0; i < order.size(); ++i) if (order[i]->nb) memcpy(nbuf + slot[i].base0, order[i]->bases, (size_t)order[i]->nb);
            free_host(B->bases, B->pinned, B->registered);
            B->bases = nbuf; B->bases_cap = total + 64; B->pinned = pin; B->registered = false;
        } else {
            for (size_t i = 0; i < order.size(); ++i) {
                const Piece& P = *order[i];
                if (P.own && P.nb) memcpy(B->bases + (P.from - first), P.bases, (size_t)P.nb);
                if (!P.ends.empty()) B->segs.push_back(Batch::Segment{slot[i].rec0, slot[i].base0, (uint64_t)(P.from - first)});
            }
        }
    } else B->reserve_bases(total + 16);
    B->offsets.resize(n_rec + 1); B->offsets[0] = 0; B->name_off.resize(n_rec); B->comment_off.resize(n_rec); B->qual_off.resize(n_rec);
    B->name_pool.resize(n_name); B->text_pool.resize((size_t)n_text);
    if (carried_text) {
        char* t = B->text_pool.data();
        if (cin->has_comment) { memcpy(t, cin->comment.c_str(), cin->comment.size() + 1); t += cin->comment.size() + 1; }
        if (cin->has_qual) memcpy(t, cin->qual.c_str(), cin->qual.size() + 1);
    }
    const double t3 = now();
    std::atomic<size_t> nextc{0};
    auto splice = [&]() {
        for (;;) {
            const size_t i = nextc.fetch_add(1);
            if (i >= order.size()) return;
            const Piece& P = *order[i];
            const Slot& S = slot[i];
            if (P.nb && !in_place) memcpy(B->bases + S.base0, P.bases, (size_t)P.nb);
            if (!P.name_pool.empty()) memcpy(B->name_pool.data() + S.name0, P.name_pool.data(), P.name_pool.size());
            if (P.text_size() && (!in_place || P.text_diverted)) memcpy(B->text_pool.data() + S.text0, P.text_data(), P.text_size());
            const size_t m = P.ends.size();
            for (size_t r = 0; r < m; ++r) {
                B->name_off[S.rec0 + r] = S.name0 + P.name_off[r];
                B->comment_off[S.rec0 + r] = P.comment_off[r] < 0 ? S.in_comment : S.text0 + P.comment_off[r];
                B->qual_off[S.rec0 + r] = P.qual_off[r] < 0 ? S.in_qual : S.text0 + P.qual_off[r];
                B->offsets[S.rec0 + r + 1] = S.base0 + P.ends[r];
            }
            if (!arena || P.own) const_cast<Piece&>(P).release();   // the per-piece vectors, on this thread (an arena keeps its pieces' memory)
        }
    };
    {
        std::vector<std::thread> pool;
        const size_t nt = n_threads < order.size() ? n_threads : order.size();
        for (size_t t = 1; t < nt; ++t) pool.emplace_back(splice);
        splice();
        for (auto& t : pool) t.join();
    }
    B->parse_status = status;
    if (cout) {
        cout->status = status;
        cout->next_hp = end_hp;
        cout->has_comment = cc >= 0; cout->has_qual = cq >= 0;
        cout->comment = cc >= 0 ? std::string(B->text_pool.data() + cc) : std::string();
        cout->qual = cq >= 0 ? std::string(B->text_pool.data() + cq) : std::string();
    }
    if (getenv("CRASS_B200_TRACE"))
        fprintf(stderr, "[crass_b200] parse: bytes [%zu, %zu): %zu pieces guessed, %zu kept in order, %zu re-parsed gaps, %zu threads; "
                "parse %.1f ms, gaps %.1f ms, alloc %.1f ms, splice %.1f ms\n",
                first, end_hp == (size_t)-1 ? n : end_hp, np, order.size() - patches.size(), patches.size(), n_threads, t1 - t0, t2 - t1, t3 - t2, now() - t3);
}
}  // namespace

int parse_file(const char* path, Batch** out, Batch* reuse) {
    Input in;
    if (!in.open(path)) return fail(CRASS_B200_EIO, std::string("cannot open ") + path);
    Batch* B = reuse ? reuse : new Batch();
    try {
        parse_range(View{in.data, in.size, in.read_error.load()}, (size_t)-1, (size_t)-1, nullptr, B, nullptr);
    } catch (std::exception& ex) {
        if (!reuse) delete B;
        return fail(CRASS_B200_ENOMEM, std::string("parse_file: ") + ex.what());
    }
    *out = B;
    return 0;
}

// ---- the streamed feed: one input parsed range by range (engine.cu overlaps the ranges with the copy, K1 and the replay) ------
struct ParseStream {
    Input in;
    size_t range_bytes = 0;
    RangeCarry carry;
    ParseArena arena;                             // scratch and piece memory, kept from range to range
    bool started = false, ended = false;
};

ParseStream* parse_stream_open(const char* path, size_t range_bytes) {
    std::unique_ptr<ParseStream> s(new ParseStream());
    if (!s->in.open_streaming(path)) { fail(CRASS_B200_EIO, std::string("cannot open ") + path); return nullptr; }
    s->range_bytes = range_bytes;
    return s.release();
}

// the input's size; for a gz archive that is still being inflated, a guess (four times the archive)
size_t parse_stream_size(const ParseStream* s) { return s->in.incremental ? (s->in.expect_size ? s->in.expect_size : s->in.map_len / 64 * 4) : s->in.size; }

// parses the next range into `reuse`; 1: a range was parsed (it may hold no record), 0: the stream had ended before, < 0: the error code
int parse_stream_next(ParseStream* s, Batch* reuse) {
    if (s->ended) return 0;
    try {
        const size_t start = s->started ? s->carry.next_hp : (size_t)-1;
        const size_t first = s->started ? start : 0;
        RangeCarry out;
        // a range may look a little past its end (the record that straddles it): for an input that is still being inflated
        // that margin must be there before the range is parsed, and if a record turns out longer than the margin the range is
        // parsed again once more has arrived
        size_t margin = env_size_early("CRASS_B200_GZ_STREAM_MARGIN", std::max<size_t>((size_t)32 << 20, s->range_bytes >> 2));
        for (;;) {
            bool final_size = true;
            const size_t have = s->in.wait_for(first + s->range_bytes + margin, &final_size);
            if (s->in.inflate_failed.load()) throw std::runtime_error("gz stream could not be inflated");
            const bool to_end = !s->range_bytes || (final_size && first + s->range_bytes + (s->range_bytes >> 2) >= have);
            parse_range(View{s->in.data, have, final_size && s->in.read_error.load(std::memory_order_acquire)}, start, to_end ? (size_t)-1 : first + s->range_bytes, s->started ? &s->carry : nullptr, reuse, &out, &s->arena);
            if (final_size || (out.status == 0 && out.next_hp != (size_t)-1)) break;
            margin *= 4;                                              // the view ended inside a record: wait for more of it
        }
        s->started = true;
        s->carry = out;
        if (out.status != 0 || out.next_hp == (size_t)-1) s->ended = true;
        if (!s->ended) reuse->parse_status = 0;               // the loop goes on in the next range
    } catch (std::exception& ex) {
        s->ended = true;
        return fail(CRASS_B200_ENOMEM, std::string("parse_stream_next: ") + ex.what());   // (the codes are negative)
    }
    return 1;
}

void parse_stream_close(ParseStream* s) { delete s; }

}  // namespace cbh
