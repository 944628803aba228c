// bed_reader.cpp -- host-side reader of fragments / regions BED files: the data
// format in front of cgs_fragment_counts_*.  Replaces, for the columns the path reads, the reference's
// read_fragments_to_pyranges (src/pycisTopic/fragments.py:21-165: skip the leading blank / '#' lines, take the
// column count from the first record, Chromosome and Name as dictionary-coded strings, Start / End as int32)
// and pyranges.read_bed at cistopic_class.py:835,578; cgs_bed_collapse_duplicates is utils.py:284-293.
//
// No device work here: the calling thread reads the next block of text -- plain files directly, BGZF (bgzip) files
// with their members inflated side by side on all threads, any other gzip file through zlib's sequential reader --
// while a pool of worker threads tokenises the previous block; every worker codes the strings it meets against its
// own dictionary and the dictionaries are merged (sorted, as a pandas "category" column is) once the file has been
// read.  The arrays a worker fills become chunks of the table as they are; the export copies the chunks side by side.
#include "../../include/cgs_b200.h"

#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <new>
#include <exception>
#include <future>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace {

// strings of one worker, coded in order of first appearance
struct LocalDict {
    std::unordered_map<std::string_view, int32_t> ids;
    std::vector<std::unique_ptr<char[]>> slabs;  // string storage that never moves (the map keys point into it)
    size_t slab_used = 0, slab_size = 0;
    std::vector<std::string_view> strings;
    // the last two lookups: consecutive records of a position-sorted file repeat the chromosome, and often the barcode
    std::string_view last_key[2];
    int32_t last_id[2] = {-1, -1};

    int32_t code(std::string_view s) {
        if (last_id[0] >= 0 && s == last_key[0]) return last_id[0];
        if (last_id[1] >= 0 && s == last_key[1]) {
            std::swap(last_key[0], last_key[1]);
            std::swap(last_id[0], last_id[1]);
            return last_id[0];
        }
        auto it = ids.find(s);
        int32_t id;
        std::string_view stored;
        if (it != ids.end()) {
            id = it->second;
            stored = it->first;
        } else {
            if (slab_used + s.size() > slab_size) {
                slab_size = std::max<size_t>(1 << 20, s.size());
                slabs.emplace_back(new char[slab_size]);
                slab_used = 0;
            }
            char *dst = slabs.back().get() + slab_used;
            memcpy(dst, s.data(), s.size());
            slab_used += s.size();
            stored = std::string_view(dst, s.size());
            id = (int32_t)strings.size();
            strings.push_back(stored);
            ids.emplace(stored, id);
        }
        last_key[1] = last_key[0];
        last_id[1] = last_id[0];
        last_key[0] = stored;
        last_id[0] = id;
        return id;
    }
};

struct Worker {
    LocalDict chrom, name;
};

struct Columns {
    std::vector<int32_t> chrom, start, end, name, score;
};

// what one worker made of one piece of a block
struct Piece : Columns {
    int64_t lines = 0;       // physical lines of the piece (for error messages)
    int64_t error_line = 0;  // 1-based inside the piece, 0 = none
    std::string error;
    bool score_not_integer = false;
};

inline bool parse_i32(const char *b, const char *e, int32_t *out) {
    if (b == e) return false;
    bool neg = false;
    if (*b == '-' || *b == '+') {
        neg = *b == '-';
        if (++b == e) return false;
    }
    int64_t v = 0;
    for (; b < e; ++b) {
        unsigned d = (unsigned)(*b - '0');
        if (d > 9) return false;
        v = v * 10 + d;
        if (v > (int64_t)1 << 31) return false;
    }
    v = neg ? -v : v;
    if (v > INT32_MAX || v < INT32_MIN) return false;
    *out = (int32_t)v;
    return true;
}

void parse_piece(const char *b, const char *e, int32_t n_columns, Worker *w, Piece *p) {
    const bool want_name = n_columns >= 4, want_score = n_columns >= 5;
    // the piece's arrays become part of the table as they are: sized once, by its number of lines
    const size_t n_lines = (size_t)std::count(b, e, '\n') + 1;  // + 1: a last line without a line end
    p->chrom.reserve(n_lines);
    p->start.reserve(n_lines);
    p->end.reserve(n_lines);
    if (want_name) p->name.reserve(n_lines);
    if (want_score) p->score.reserve(n_lines);
    while (b < e) {
        const char *nl = (const char *)memchr(b, '\n', (size_t)(e - b));
        const char *le = nl ? nl : e;
        const char *next = nl ? nl + 1 : e;
        ++p->lines;
        if (le > b && le[-1] == '\r') --le;
        if (le == b) {  // blank line (the pandas / polars readers skip them)
            b = next;
            continue;
        }
        // fields
        const char *f[5];
        const char *g[5];
        int32_t nf = 0;
        const char *c = b;
        for (;;) {
            const char *tab = (const char *)memchr(c, '\t', (size_t)(le - c));
            const char *fe = tab ? tab : le;
            if (nf < 5) {
                f[nf] = c;
                g[nf] = fe;
            }
            ++nf;
            if (!tab) break;
            c = tab + 1;
        }
        if (nf != n_columns) {
            p->error_line = p->lines;
            p->error = "expected " + std::to_string(n_columns) + " columns, saw " + std::to_string(nf);
            return;
        }
        int32_t s, t;
        if (!parse_i32(f[1], g[1], &s) || !parse_i32(f[2], g[2], &t)) {
            p->error_line = p->lines;
            p->error = "Start / End is not a 32-bit integer";
            return;
        }
        p->chrom.push_back(w->chrom.code(std::string_view(f[0], (size_t)(g[0] - f[0]))));
        p->start.push_back(s);
        p->end.push_back(t);
        if (want_name) p->name.push_back(w->name.code(std::string_view(f[3], (size_t)(g[3] - f[3]))));
        if (want_score) {
            int32_t v = 0;
            if (!parse_i32(f[4], g[4], &v)) p->score_not_integer = true;
            p->score.push_back(v);
        }
        b = next;
    }
}

// fn(0) .. fn(n - 1) side by side; the first exception of any of them is rethrown here once all have ended
void run_parallel(int n, const std::function<void(int)> &fn) {
    std::vector<std::thread> th;
    std::exception_ptr first;
    std::mutex mu;
    auto guarded = [&](int i) {
        try {
            fn(i);
        } catch (...) {
            std::lock_guard<std::mutex> lock(mu);
            if (!first) first = std::current_exception();
        }
    };
    for (int i = 1; i < n; ++i) {
        try {
            th.emplace_back(guarded, i);
        } catch (...) {  // no thread to be had: that share is done here
            guarded(i);
        }
    }
    guarded(0);
    for (auto &t : th) t.join();
    if (first) std::rethrow_exception(first);
}

// ---- the file as a stream of text ------------------------------------------------------------------------------
// plain text is read as it is; a BGZF file (bgzip: what CellRanger and tabix-indexed fragments files are -- a series of
// gzip members of at most 64 KiB that name their own size) is inflated by all threads, member by member; any other gzip
// file goes through zlib's sequential reader.
struct Source {
    FILE *f = nullptr;
    gzFile gz = nullptr;
    bool bgzf = false;
    int threads = 1;
    std::vector<unsigned char> raw;  // BGZF: compressed bytes not yet inflated
    size_t raw_pos = 0;
    bool raw_eof = false;
    std::vector<char> spill;  // BGZF: inflated bytes that did not fit the caller's buffer
    size_t spill_pos = 0;

    ~Source() {
        if (gz) gzclose(gz);
        if (f) fclose(f);
    }

    // size of the BGZF member that starts at p (n bytes available): 0 = need more bytes, -1 = not a BGZF member
    static long member_size(const unsigned char *p, size_t n) {
        if (n < 18) return 0;
        if (p[0] != 0x1f || p[1] != 0x8b || p[2] != 8 || !(p[3] & 4)) return -1;
        size_t xlen = (size_t)p[10] | ((size_t)p[11] << 8);
        if (n < 12 + xlen) return 0;
        for (size_t o = 12; o + 4 <= 12 + xlen;) {
            size_t len = (size_t)p[o + 2] | ((size_t)p[o + 3] << 8);
            if (p[o] == 'B' && p[o + 1] == 'C' && len == 2 && o + 6 <= 12 + xlen)
                return (long)((size_t)p[o + 4] | ((size_t)p[o + 5] << 8)) + 1;
            o += 4 + len;
        }
        return -1;
    }

    bool open(const char *path, int n_threads, std::string *err) {
        threads = n_threads;
        f = fopen(path, "rb");
        if (!f) {
            *err = std::string("cannot open ") + path;
            return false;
        }
        unsigned char head[512];
        size_t n = fread(head, 1, sizeof head, f);
        const bool gzip = n >= 2 && head[0] == 0x1f && head[1] == 0x8b;
        bgzf = gzip && member_size(head, n) > 0;
        if (gzip && !bgzf) {
            fclose(f);
            f = nullptr;
            gz = gzopen(path, "rb");
            if (!gz) {
                *err = std::string("cannot open ") + path;
                return false;
            }
            gzbuffer(gz, 1 << 20);
            return true;
        }
        if (fseek(f, 0, SEEK_SET) != 0) {
            *err = std::string("cannot rewind ") + path;
            return false;
        }
        return true;
    }

    // fills up to n bytes; fewer than n only at the end of the stream; -1 on error
    int64_t read(char *dst, size_t n, std::string *err) {
        if (bgzf) return read_bgzf(dst, n, err);
        size_t got = 0;
        if (f) {
            got = fread(dst, 1, n, f);
            if (got < n && ferror(f)) {
                *err = "read error";
                return -1;
            }
            return (int64_t)got;
        }
        while (got < n) {
            unsigned want = (unsigned)std::min<size_t>(n - got, 1u << 30);
            int r = gzread(gz, dst + got, want);
            int code = 0;
            if (r <= 0) {  // 0 is the end -- unless the stream was cut short, which zlib tells through gzerror only
                const char *msg = gzerror(gz, &code);
                if (r < 0 || (code != Z_OK && code != Z_STREAM_END)) {
                    *err = msg && *msg ? msg : "damaged gzip stream";
                    return -1;
                }
                break;
            }
            got += (size_t)r;
        }
        return (int64_t)got;
    }

    struct Member {
        size_t offset, size, out_offset, out_size;
    };

    static bool inflate_member(const unsigned char *p, size_t size, char *dst, size_t out_size, std::string *err) {
        const size_t xlen = (size_t)p[10] | ((size_t)p[11] << 8), header = 12 + xlen;
        if (size < header + 8) {
            *err = "BGZF member shorter than its header";
            return false;
        }
        z_stream z;
        memset(&z, 0, sizeof z);
        if (inflateInit2(&z, -15) != Z_OK) {
            *err = "inflateInit2 failed";
            return false;
        }
        z.next_in = const_cast<unsigned char *>(p + header);
        z.avail_in = (uInt)(size - header - 8);
        z.next_out = (unsigned char *)dst;
        z.avail_out = (uInt)out_size;
        int rc = inflate(&z, Z_FINISH);
        size_t produced = z.total_out;
        inflateEnd(&z);
        const unsigned char *t = p + size - 8;
        uint32_t crc = (uint32_t)t[0] | ((uint32_t)t[1] << 8) | ((uint32_t)t[2] << 16) | ((uint32_t)t[3] << 24);
        if (rc != Z_STREAM_END || produced != out_size ||
            (uint32_t)crc32(crc32(0L, Z_NULL, 0), (const unsigned char *)dst, (uInt)out_size) != crc) {
            *err = "damaged BGZF member (inflate / size / CRC check)";
            return false;
        }
        return true;
    }

    int64_t read_bgzf(char *dst, size_t n, std::string *err) {
        size_t got = 0;
        for (;;) {
            if (spill_pos < spill.size()) {  // what an earlier call could not place
                size_t take = std::min(n - got, spill.size() - spill_pos);
                memcpy(dst + got, spill.data() + spill_pos, take);
                spill_pos += take;
                got += take;
            }
            if (got == n) return (int64_t)got;
            // compressed bytes: keep about as many as the caller wants text, refill when a member is incomplete
            if (!raw_eof && raw.size() - raw_pos < (size_t)1 << 17) {
                raw.erase(raw.begin(), raw.begin() + (ptrdiff_t)raw_pos);
                raw_pos = 0;
                size_t want = std::max<size_t>((size_t)1 << 20, std::min<size_t>(n - got, (size_t)64 << 20));
                size_t old = raw.size();
                raw.resize(old + want);
                size_t r = fread(raw.data() + old, 1, want, f);
                if (r < want && ferror(f)) {
                    *err = "read error";
                    return -1;
                }
                raw.resize(old + r);
                if (r < want) raw_eof = true;
            }
            if (raw_pos == raw.size()) {
                if (raw_eof) return (int64_t)got;
                continue;
            }
            // the complete members in the buffer that fit what is left of dst
            std::vector<Member> batch;
            size_t p = raw_pos, out = 0;
            bool first_too_big = false;
            while (p < raw.size()) {
                long ms = member_size(raw.data() + p, raw.size() - p);
                if (ms < 0) {
                    *err = "not a BGZF member where one is expected (a file that mixes bgzip and gzip output?)";
                    return -1;
                }
                if (ms == 0 || p + (size_t)ms > raw.size()) {
                    if (raw_eof) {
                        *err = "the file ends inside a BGZF member";
                        return -1;
                    }
                    break;
                }
                const unsigned char *t = raw.data() + p + ms - 4;
                size_t isize = (size_t)t[0] | ((size_t)t[1] << 8) | ((size_t)t[2] << 16) | ((size_t)t[3] << 24);
                if (isize > 65536) {  // the format's limit; a damaged trailer must not size a buffer
                    *err = "BGZF member claims more than 64 KiB of text";
                    return -1;
                }
                if (got + out + isize > n) {
                    first_too_big = batch.empty();
                    break;
                }
                batch.push_back({p, (size_t)ms, out, isize});
                out += isize;
                p += (size_t)ms;
            }
            if (first_too_big) {  // one member into the spill buffer, to be handed out in parts
                long ms = member_size(raw.data() + raw_pos, raw.size() - raw_pos);
                const unsigned char *t = raw.data() + raw_pos + ms - 4;
                size_t isize = (size_t)t[0] | ((size_t)t[1] << 8) | ((size_t)t[2] << 16) | ((size_t)t[3] << 24);
                spill.resize(isize);
                spill_pos = 0;
                if (!inflate_member(raw.data() + raw_pos, (size_t)ms, spill.data(), isize, err)) return -1;
                raw_pos += (size_t)ms;
                continue;
            }
            if (batch.empty()) {
                if (p + 18 > raw.size() && raw_eof && p < raw.size()) {
                    *err = "the file ends inside a BGZF member";
                    return -1;
                }
                // an incomplete member at the end of the buffer: read on (the refill above triggers below 128 KiB)
                if (raw.size() - raw_pos >= (size_t)1 << 17) {
                    *err = "BGZF member larger than 64 KiB";
                    return -1;
                }
                continue;
            }
            std::atomic<size_t> next{0};
            std::atomic<bool> ok{true};
            std::vector<std::string> errors((size_t)threads);
            char *base = dst + got;
            run_parallel(std::min<int>(threads, (int)batch.size()), [&](int t) {
                for (;;) {
                    size_t i = next.fetch_add(1);
                    if (i >= batch.size() || !ok.load()) return;
                    const Member &m = batch[i];
                    if (!inflate_member(raw.data() + m.offset, m.size, base + m.out_offset, m.out_size, &errors[(size_t)t]))
                        ok.store(false);
                }
            });
            if (!ok.load()) {
                for (auto &e : errors)
                    if (!e.empty()) *err = e;
                return -1;
            }
            got += out;
            raw_pos = p;
            if (got == n) return (int64_t)got;
        }
    }
};

struct Segment {
    size_t chunk;
    int32_t worker;
};

}  // namespace

struct cgs_bed_table {
    int32_t n_columns = 0;
    int32_t score_state = 0;  // 0: no fifth column, 1: integers, 2: a fifth column that is not all integers
    int threads = 1;
    std::vector<Columns> chunks;  // the records, in file order: chunk after chunk (one chunk when flat)
    std::vector<std::string> chrom_names, name_names;  // sorted
    size_t rows() const {
        size_t n = 0;
        for (auto &c : chunks) n += c.chrom.size();
        return n;
    }
};

namespace {

// sorted union of the workers' strings and, per worker, local code -> position in it
void merge_dicts(std::vector<Worker> &workers, bool names, std::vector<std::string> *sorted,
                 std::vector<std::vector<int32_t>> *remap) {
    std::vector<std::string_view> all;
    for (auto &w : workers) {
        auto &d = names ? w.name : w.chrom;
        all.insert(all.end(), d.strings.begin(), d.strings.end());
    }
    std::sort(all.begin(), all.end());
    all.erase(std::unique(all.begin(), all.end()), all.end());
    sorted->assign(all.begin(), all.end());
    remap->resize(workers.size());
    run_parallel((int)workers.size(), [&](int i) {
        auto &d = names ? workers[(size_t)i].name : workers[(size_t)i].chrom;
        auto &r = (*remap)[(size_t)i];
        r.resize(d.strings.size());
        for (size_t j = 0; j < d.strings.size(); ++j)
            r[j] = (int32_t)(std::lower_bound(all.begin(), all.end(), d.strings[j]) - all.begin());
    });
}

// all chunks into one (what the de-duplication works on)
void flatten(cgs_bed_table *t) {
    if (t->chunks.size() <= 1) return;
    const size_t n = t->rows();
    Columns flat;
    auto gather = [&](std::vector<int32_t> Columns::*col) {
        auto &dst = flat.*col;
        if ((t->chunks[0].*col).empty() && n > 0) {
            bool any = false;
            for (auto &c : t->chunks) any |= !(c.*col).empty();
            if (!any) return;
        }
        dst.reserve(n);
        for (auto &c : t->chunks) {
            dst.insert(dst.end(), (c.*col).begin(), (c.*col).end());
            std::vector<int32_t>().swap(c.*col);
        }
    };
    gather(&Columns::chrom);
    gather(&Columns::start);
    gather(&Columns::end);
    gather(&Columns::name);
    gather(&Columns::score);
    t->chunks.clear();
    t->chunks.push_back(std::move(flat));
}

}  // namespace

// C++ exceptions (allocation failure, no thread to be had) end at the ABI as an error code
#define CGS_BED_GUARD(call)                                                            \
    try {                                                                              \
        return call;                                                                   \
    } catch (const std::bad_alloc &) {                                                 \
        CGS_FAIL("out of host memory");                                                \
    } catch (const std::exception &e) {                                                \
        CGS_FAIL(std::string("unexpected failure: ") + e.what());                      \
    }

static int bed_read(cgs_bed_table **out, const char *path, int32_t min_columns, int32_t n_threads) {
    if (!out || !path) CGS_FAIL("NULL argument");
    if (min_columns < 3 || min_columns > 5) CGS_FAIL("min_columns must be 3, 4 or 5");
    *out = nullptr;
    int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    T = std::max(1, std::min(T, 64));
    std::string err;
    Source in;
    if (!in.open(path, T, &err)) CGS_FAIL(err);

    size_t kBlock = (size_t)32 << 20;
    if (const char *v = getenv("CGS_BED_BLOCK_BYTES")) {  // tests shrink the block to walk every carry-over case
        long long n = atoll(v);
        if (n >= 16) kBlock = (size_t)n;
    }
    // two text buffers (never zero-filled: the pages are touched once, by the read)
    std::unique_ptr<char[]> buf[2];
    size_t cap[2] = {0, 0};
    auto ensure = [&](int which, size_t bytes, size_t keep) {
        if (cap[which] >= bytes) return;
        std::unique_ptr<char[]> grown(new char[bytes]);
        if (keep) memcpy(grown.get(), buf[which].get(), keep);
        buf[which] = std::move(grown);
        cap[which] = bytes;
    };
    size_t have = 0;  // bytes of buf[cur] in use
    int cur = 0;
    bool eof = false;
    int64_t lines_before = 0;  // physical lines consumed so far

    auto fill = [&](int which, size_t keep) -> int {  // read behind `keep` carried bytes, one block
        ensure(which, keep + kBlock, keep);
        int64_t r = in.read(buf[which].get() + keep, kBlock, &err);
        if (r < 0) return 1;
        have = keep + (size_t)r;
        if ((size_t)r < kBlock) eof = true;
        return 0;
    };
    auto io_error = [&]() { return ::cgs::fail(__FILE__, __LINE__, std::string(path) + ": " + err); };
    if (fill(cur, 0)) return io_error();

    // ---- the header: blank lines and '#' lines before the first record; its column count (fragments.py:67-91)
    size_t pos = 0;
    int32_t n_columns = 0;
    for (;;) {
        if (pos >= have) {
            if (eof) break;
            if (fill(cur, 0)) return io_error();  // everything so far was skipped
            pos = 0;
            continue;
        }
        const char *b = buf[cur].get() + pos;
        const char *nl = (const char *)memchr(b, '\n', have - pos);
        if (!nl && !eof) {  // the line continues behind the buffer
            memmove(buf[cur].get(), b, have - pos);
            size_t keep = have - pos;
            pos = 0;
            if (fill(cur, keep)) return io_error();
            continue;
        }
        const char *le = nl ? nl : buf[cur].get() + have;
        const char *sb = b, *se = le;  // strip() of the reference: white space on both sides
        while (sb < se && (*sb == ' ' || *sb == '\t' || *sb == '\r')) ++sb;
        while (se > sb && (se[-1] == ' ' || se[-1] == '\t' || se[-1] == '\r')) --se;
        if (sb == se || *sb == '#') {
            pos = (size_t)((nl ? nl + 1 : le) - buf[cur].get());
            ++lines_before;
            continue;
        }
        n_columns = 1 + (int32_t)std::count(sb, se, '\t');
        break;
    }
    if (n_columns < min_columns) {
        // the reference's message (fragments.py:87-91) for a fragments file; the same wording for a regions file
        CGS_FAIL(std::string(min_columns >= 4 ? "Fragments BED file" : "BED file") + " needs to have at least " +
                 std::to_string(min_columns) + " columns. \"" + path + "\" contains only " + std::to_string(n_columns) +
                 " columns.");
    }

    auto table = std::make_unique<cgs_bed_table>();
    table->n_columns = n_columns;
    table->threads = T;
    std::vector<Worker> workers((size_t)T);
    std::vector<Segment> segments;
    bool score_not_integer = false;

    // ---- blocks: block i is tokenised by the pool while block i + 1 is read (and inflated) by this thread
    std::vector<Piece> pieces;
    std::future<void> pending;  // declared after everything the pool touches: its destructor waits for the pool
    auto launch = [&](const char *b, const char *e) {
        pieces.assign((size_t)T, Piece());
        pending = std::async(std::launch::async, [&, b, e]() {
            std::vector<const char *> cut((size_t)T + 1);  // [b, e) cut at line ends into T pieces
            cut[0] = b;
            cut[(size_t)T] = e;
            for (int i = 1; i < T; ++i) {
                const char *c = std::max(cut[(size_t)i - 1], b + (size_t)(e - b) * (size_t)i / (size_t)T);
                const char *nl = c < e ? (const char *)memchr(c, '\n', (size_t)(e - c)) : nullptr;
                cut[(size_t)i] = nl ? nl + 1 : e;
            }
            run_parallel(T, [&](int i) {
                parse_piece(cut[(size_t)i], cut[(size_t)i + 1], n_columns, &workers[(size_t)i], &pieces[(size_t)i]);
            });
        });
    };
    auto collect = [&]() -> int {  // the pool's pieces, in file order, become chunks of the table
        pending.get();
        for (int i = 0; i < T; ++i) {
            Piece &p = pieces[(size_t)i];
            if (p.error_line) {
                return ::cgs::fail(__FILE__, __LINE__,
                                   std::string(path) + ": line " + std::to_string(lines_before + p.error_line) + ": " + p.error);
            }
            lines_before += p.lines;
            if (p.chrom.empty()) continue;
            score_not_integer |= p.score_not_integer;
            segments.push_back({table->chunks.size(), i});
            table->chunks.push_back(std::move(static_cast<Columns &>(p)));
        }
        return 0;
    };

    bool running = false;  // the pool works on the OTHER buffer
    for (;;) {
        char *base = buf[cur].get();
        size_t end = have;  // [pos, end): whole lines (everything that is left at the end of the stream)
        if (!eof) {
            while (end > pos && base[end - 1] != '\n') --end;
            if (end == pos) {  // a line longer than a block: read on behind it
                memmove(base, base + pos, have - pos);
                size_t keep = have - pos;
                pos = 0;
                if (fill(cur, keep)) {
                    if (running) pending.get();
                    return io_error();
                }
                continue;
            }
        }
        if (running) {  // the other buffer is about to be overwritten
            running = false;
            if (collect()) return 1;
        }
        launch(base + pos, base + end);
        running = true;
        if (eof) break;
        const int nxt = cur ^ 1;
        const size_t carry = have - end;  // the unfinished last line travels to the next block
        ensure(nxt, carry + kBlock, 0);
        memcpy(buf[nxt].get(), base + end, carry);
        cur = nxt;
        pos = 0;
        if (fill(cur, carry)) {
            pending.get();
            return io_error();
        }
    }
    if (running && collect()) return 1;

    // ---- one dictionary per column, sorted; records recoded
    std::vector<std::vector<int32_t>> remap_c, remap_n;
    merge_dicts(workers, false, &table->chrom_names, &remap_c);
    const bool has_name = n_columns >= 4;
    if (has_name) merge_dicts(workers, true, &table->name_names, &remap_n);
    std::atomic<size_t> next_seg{0};
    run_parallel(T, [&](int) {
        for (;;) {
            size_t s = next_seg.fetch_add(1);
            if (s >= segments.size()) return;
            Columns &c = table->chunks[segments[s].chunk];
            const auto &rc = remap_c[(size_t)segments[s].worker];
            for (auto &v : c.chrom) v = rc[(size_t)v];
            if (has_name) {
                const auto &rn = remap_n[(size_t)segments[s].worker];
                for (auto &v : c.name) v = rn[(size_t)v];
            }
        }
    });
    table->score_state = n_columns >= 5 ? (score_not_integer ? 2 : 1) : 0;
    *out = table.release();
    return 0;
}

extern "C" {

int cgs_bed_read(cgs_bed_table **out, const char *path, int32_t min_columns, int32_t n_threads) {
    CGS_BED_GUARD(bed_read(out, path, min_columns, n_threads))
}

int cgs_bed_destroy(cgs_bed_table *t) {
    delete t;
    return 0;
}

static int64_t chars_of(const std::vector<std::string> &v) {
    int64_t n = 0;
    for (auto &s : v) n += (int64_t)s.size();
    return n;
}

int cgs_bed_dims(const cgs_bed_table *t, int64_t *n_rows, int32_t *n_columns, int32_t *n_chromosomes,
                 int64_t *chromosome_chars, int32_t *n_names, int64_t *name_chars, int32_t *score_state) {
    if (!t) CGS_FAIL("NULL table");
    if (n_rows) *n_rows = (int64_t)t->rows();
    if (n_columns) *n_columns = t->n_columns;
    if (n_chromosomes) *n_chromosomes = (int32_t)t->chrom_names.size();
    if (chromosome_chars) *chromosome_chars = chars_of(t->chrom_names);
    if (n_names) *n_names = (int32_t)t->name_names.size();
    if (name_chars) *name_chars = chars_of(t->name_names);
    if (score_state) *score_state = t->score_state;
    return 0;
}

static void export_strings(const std::vector<std::string> &v, char *chars, int64_t *offsets) {
    int64_t o = 0;
    for (size_t i = 0; i < v.size(); ++i) {
        if (offsets) offsets[i] = o;
        if (chars) memcpy(chars + o, v[i].data(), v[i].size());
        o += (int64_t)v[i].size();
    }
    if (offsets) offsets[v.size()] = o;
}

static int bed_export(const cgs_bed_table *t, int32_t *chromosome, int32_t *start, int32_t *end, int32_t *name, int32_t *score,
                      char *chromosome_chars, int64_t *chromosome_offsets, char *name_chars, int64_t *name_offsets) {
    if (!t) CGS_FAIL("NULL table");
    if (name && t->n_columns < 4) CGS_FAIL("the file has no Name column");
    if (score && t->score_state == 0) CGS_FAIL("the table has no Score column");
    std::vector<size_t> first(t->chunks.size() + 1, 0);
    for (size_t i = 0; i < t->chunks.size(); ++i) first[i + 1] = first[i] + t->chunks[i].chrom.size();
    std::atomic<size_t> next{0};
    run_parallel(std::max(1, std::min<int>(t->threads, (int)t->chunks.size())), [&](int) {  // the chunks side by side
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= t->chunks.size()) return;
            const Columns &c = t->chunks[i];
            const size_t bytes = c.chrom.size() * sizeof(int32_t);
            if (chromosome) memcpy(chromosome + first[i], c.chrom.data(), bytes);
            if (start) memcpy(start + first[i], c.start.data(), bytes);
            if (end) memcpy(end + first[i], c.end.data(), bytes);
            if (name) memcpy(name + first[i], c.name.data(), bytes);
            if (score) memcpy(score + first[i], c.score.data(), bytes);
        }
    });
    export_strings(t->chrom_names, chromosome_chars, chromosome_offsets);
    if (t->n_columns >= 4) export_strings(t->name_names, name_chars, name_offsets);
    return 0;
}

int cgs_bed_export(const cgs_bed_table *t, int32_t *chromosome, int32_t *start, int32_t *end, int32_t *name, int32_t *score,
                   char *chromosome_chars, int64_t *chromosome_offsets, char *name_chars, int64_t *name_offsets) {
    CGS_BED_GUARD(bed_export(t, chromosome, start, end, name, score, chromosome_chars, chromosome_offsets, name_chars,
                             name_offsets))
}

// utils.py:284-293 collapse_duplicates: records equal in (Chromosome, Start, End, Name) become one record whose Score is
// their number.  The first record of every group stays where it is (the reference returns the groups sorted; the order of
// the records does not reach the count matrix).
static int bed_collapse_duplicates(cgs_bed_table *t, int32_t n_threads, int64_t *n_removed) {
    if (!t) CGS_FAIL("NULL table");
    if (t->n_columns < 4) CGS_FAIL("collapsing duplicates needs the Name column");
    flatten(t);
    if (t->chunks.empty()) t->chunks.emplace_back();
    Columns &f = t->chunks[0];
    const size_t n = f.chrom.size();
    int T = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    T = std::max(1, std::min(T, 64));
    if (n < (size_t)1 << 16) T = 1;
    std::vector<int32_t> count(n, 0);  // > 0 at the first record of a group
    auto mix = [](uint64_t a, uint64_t b) {
        uint64_t h = (a ^ (b * 0x9E3779B97F4A7C15ull)) * 0xBF58476D1CE4E5B9ull;
        h ^= h >> 31;
        h *= 0x94D049BB133111EBull;
        return h ^ (h >> 29);
    };
    const int32_t *c = f.chrom.data(), *s = f.start.data(), *e = f.end.data(), *b = f.name.data();
    auto key_a = [&](size_t i) { return ((uint64_t)(uint32_t)c[i] << 32) | (uint32_t)s[i]; };
    auto key_b = [&](size_t i) { return ((uint64_t)(uint32_t)e[i] << 32) | (uint32_t)b[i]; };
    // every thread owns the keys whose hash falls into its part: it scans all records and keeps a table of its own
    run_parallel(T, [&](int part) {
        size_t mine = 0;
        for (size_t i = 0; i < n; ++i) mine += (mix(key_a(i), key_b(i)) >> 40) % (uint64_t)T == (uint64_t)part;
        size_t cap = 16;
        while (cap < 2 * mine + 2) cap <<= 1;
        std::vector<int64_t> slot(cap, -1);  // the first record with that key
        for (size_t i = 0; i < n; ++i) {
            uint64_t h = mix(key_a(i), key_b(i));
            if ((h >> 40) % (uint64_t)T != (uint64_t)part) continue;
            size_t p = (size_t)h & (cap - 1);
            for (;;) {
                int64_t r = slot[p];
                if (r < 0) {
                    slot[p] = (int64_t)i;
                    count[i] = 1;
                    break;
                }
                if (key_a((size_t)r) == key_a(i) && key_b((size_t)r) == key_b(i)) {
                    ++count[(size_t)r];
                    break;
                }
                p = (p + 1) & (cap - 1);
            }
        }
    });
    size_t o = 0;
    for (size_t i = 0; i < n; ++i) {
        if (!count[i]) continue;
        f.chrom[o] = f.chrom[i];
        f.start[o] = f.start[i];
        f.end[o] = f.end[i];
        f.name[o] = f.name[i];
        count[o] = count[i];
        ++o;
    }
    f.chrom.resize(o);
    f.start.resize(o);
    f.end.resize(o);
    f.name.resize(o);
    count.resize(o);
    f.score.swap(count);
    t->score_state = 1;
    if (n_removed) *n_removed = (int64_t)(n - o);
    return 0;
}

int cgs_bed_collapse_duplicates(cgs_bed_table *t, int32_t n_threads, int64_t *n_removed) {
    CGS_BED_GUARD(bed_collapse_duplicates(t, n_threads, n_removed))
}

}  // extern "C"
