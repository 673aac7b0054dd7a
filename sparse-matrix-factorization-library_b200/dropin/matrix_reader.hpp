/*
 * matrix_reader.hpp -- input side of the drop-in library (once the factorization takes 0.2 s, reading
 * the matrix must not take seconds).
 *
 * 1. MatrixMarket coordinate text, same semantics as the reference's SparseFrame_read_matrix_triplet (CH.c:400-525:
 *    banner "%%MatrixMarket object format field symmetry", field != "real" means complex, symmetry == "symmetric" or
 *    anything else, comment lines before the size line, entries "i j re [im]" 1-based, entries whose value is zero are
 *    dropped, nothing is summed, file order is kept) -- but the file is read in one piece and its lines are parsed by
 *    all host threads with hand-written integer / decimal scanners instead of getline + sscanf per line.
 *    The decimal scanner is exact: mantissas of at most 15 digits with a power of ten of at most 22 are one correctly
 *    rounded IEEE operation (Clinger's fast path), everything else goes through strtod, so the values are bit-identical
 *    to what "%lg" produces.
 * 2. A binary triplet container ("SFB200T1"): header + Ti + Tj + Tx in file order, so that compress() builds the very
 *    same Cp/Ci/Cx as from the text file.  SparseFrame_read_matrix recognises it by its magic; files are produced with
 *    SparseFrame_write_binary_matrix (exported by the drop-in library) or sparseframe_b200.matrices.write_binary.
 */
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

namespace sfio {

struct Triplets {
    bool symmetric = false, complex_field = false;
    long nrow = 0, ncol = 0, nz_declared = 0;
    std::vector<int64_t> ti, tj;     // 0-based
    std::vector<double> tx;          // one double per entry; (re, im) pairs when complex_field (CH.c:505-511)
};

enum { OK = 0, ERR_FORMAT = 1, ERR_ENTRY = 2, ERR_NZMAX = 3, ERR_IMAG = 4, ERR_IO = 5 };

static const char BIN_MAGIC[8] = {'S', 'F', 'B', '2', '0', '0', 'T', '1'};

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// "%ld": optional sign, decimal digits.  Returns false if no digit was read.
inline bool scan_long(const char *&p, const char *end, long &out) {
    while (p < end && is_space(*p)) ++p;
    bool neg = false;
    if (p < end && (*p == '-' || *p == '+')) { neg = *p == '-'; ++p; }
    if (p >= end || *p < '0' || *p > '9') return false;
    unsigned long v = 0;
    while (p < end && *p >= '0' && *p <= '9') { v = v * 10 + (unsigned long)(*p - '0'); ++p; }
    out = neg ? -(long)v : (long)v;
    return true;
}

// "%lg".  Fast path for plain decimals; anything unusual (hex floats, inf/nan, long mantissas, big exponents) is handed
// to strtod on a NUL-terminated copy of the token.
inline bool scan_double(const char *&p, const char *end, double &out) {
    static const double P10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                   1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    while (p < end && is_space(*p)) ++p;
    const char *start = p;
    bool neg = false;
    if (p < end && (*p == '-' || *p == '+')) { neg = *p == '-'; ++p; }
    uint64_t mant = 0;
    int sig = 0, dec_exp = 0;
    bool any = false, fast = true;
    while (p < end && *p >= '0' && *p <= '9') {
        any = true;
        if (sig > 0 || *p != '0') { if (++sig <= 19) mant = mant * 10 + (uint64_t)(*p - '0'); else ++dec_exp; }
        ++p;
    }
    if (p < end && *p == '.') {
        ++p;
        while (p < end && *p >= '0' && *p <= '9') {
            any = true;
            if (sig > 0 || *p != '0') { if (++sig <= 19) { mant = mant * 10 + (uint64_t)(*p - '0'); --dec_exp; } }
            else --dec_exp;
            ++p;
        }
    }
    if (!any) {                               // inf / nan / garbage: let strtod decide
        fast = false;
    } else if (p < end && (*p == 'e' || *p == 'E' || *p == 'd' || *p == 'D')) {
        if (*p == 'd' || *p == 'D') fast = false;          // Fortran exponents are not "%lg": strtod stops before them
        else {
            const char *q = p + 1;
            bool eneg = false;
            if (q < end && (*q == '-' || *q == '+')) { eneg = *q == '-'; ++q; }
            if (q < end && *q >= '0' && *q <= '9') {
                int e = 0;
                while (q < end && *q >= '0' && *q <= '9') { if (e < 100000) e = e * 10 + (*q - '0'); ++q; }
                dec_exp += eneg ? -e : e;
                p = q;
            }
        }
    } else if (p < end && (*p == 'x' || *p == 'X')) {
        fast = false;                          // hex float
    }
    if (fast && sig <= 15 && dec_exp >= -22 && dec_exp <= 22) {
        double v = (double)mant;               // exact: < 2^53
        if (dec_exp < 0) v /= P10[-dec_exp]; else v *= P10[dec_exp];
        out = neg ? -v : v;
        return true;
    }
    if (fast && mant == 0) { out = neg ? -0.0 : 0.0; return true; }
    // slow path: NUL-terminated copy of the token
    const char *q = start;
    while (q < end && !is_space(*q) && *q != '\n') ++q;
    char tmp[128];
    const size_t len = (size_t)(q - start);
    if (len == 0 || len >= sizeof(tmp)) return false;
    memcpy(tmp, start, len);
    tmp[len] = 0;
    char *stop = nullptr;
    out = strtod(tmp, &stop);
    if (stop == tmp) return false;
    p = start + (stop - tmp);
    return true;
}

// entries of buf[b, e): whole lines only.  Returns an error code; entries are appended to the chunk's vectors.
struct Chunk {
    std::vector<int64_t> ti, tj;
    std::vector<double> tx;
    int err = OK;
};

inline void parse_entries(const char *p, const char *end, bool complex_field, Chunk &c) {
    while (p < end) {
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        const char *le = nl ? nl : end;
        const char *q = p;
        while (q < le && is_space(*q)) ++q;
        if (q < le) {                          // blank lines are skipped
            long i, j;
            double v, w = 0.0;
            if (!scan_long(q, le, i) || !scan_long(q, le, j) || !scan_double(q, le, v)) { c.err = ERR_ENTRY; return; }
            const bool has_w = scan_double(q, le, w);
            if (complex_field && !has_w) { c.err = ERR_IMAG; return; }
            if (v != 0.0 || (has_w && w != 0.0)) {          // CH.c:493: explicit zeros are dropped
                c.ti.push_back(i - 1);
                c.tj.push_back(j - 1);
                c.tx.push_back(v);
                if (complex_field) c.tx.push_back(w);
            }
        }
        p = nl ? nl + 1 : end;
    }
}

inline int parse_matrix_market(const char *buf, size_t len, Triplets &T, int nthreads = 0) {
    const char *p = buf, *end = buf + len;
    auto line_end = [&](const char *s) { const char *nl = (const char *)memchr(s, '\n', (size_t)(end - s)); return nl ? nl : end; };
    // banner: first line (CH.c:423-445)
    if (len < 14 || strncmp(p, "%%MatrixMarket", 14) != 0) return ERR_FORMAT;
    {
        const char *le = line_end(p);
        std::vector<std::string> tok;
        const char *q = p;
        while (q < le && tok.size() < 5) {
            while (q < le && is_space(*q)) ++q;
            const char *s = q;
            while (q < le && !is_space(*q)) ++q;
            if (q > s) tok.emplace_back(s, q);
        }
        T.complex_field = !(tok.size() > 3 && tok[3] == "real");
        T.symmetric = tok.size() > 4 && tok[4] == "symmetric";
        p = le < end ? le + 1 : end;
    }
    // comments, then the size line (CH.c:447-459)
    for (;;) {
        if (p >= end) return ERR_FORMAT;
        const char *le = line_end(p);
        const char *q = p;
        while (q < le && is_space(*q)) ++q;
        if (*p == '%' || q == le) { p = le < end ? le + 1 : end; continue; }
        if (!scan_long(q, le, T.nrow) || !scan_long(q, le, T.ncol) || !scan_long(q, le, T.nz_declared)) return ERR_FORMAT;
        p = le < end ? le + 1 : end;
        break;
    }
    // entries: split at line boundaries, one chunk per thread, concatenated in file order
    const size_t body = (size_t)(end - p);
    int nt = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, (int)(body / (1u << 20)) + 1));
    std::vector<const char *> cut(nt + 1);
    cut[0] = p;
    cut[nt] = end;
    for (int t = 1; t < nt; ++t) {
        const char *s = p + body / nt * t;
        s = std::max(s, cut[t - 1]);
        const char *nl = (const char *)memchr(s, '\n', (size_t)(end - s));
        cut[t] = nl ? nl + 1 : end;
    }
    std::vector<Chunk> chunks(nt);
    // the declared count is only a hint for the reservation: a chunk cannot hold more entries than bytes / 4 ("1 1\n" ...),
    // so a damaged size line (a count of 10^15) cannot ask for an absurd allocation
    const size_t guess = std::min((size_t)std::max<long>(T.nz_declared, 0) / nt, body / nt / 4) + 16;
    auto work = [&](int t) {
        chunks[t].ti.reserve(guess); chunks[t].tj.reserve(guess); chunks[t].tx.reserve(guess);
        parse_entries(cut[t], cut[t + 1], T.complex_field, chunks[t]);
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    size_t total = 0;
    for (auto &c : chunks) { if (c.err) return c.err; total += c.ti.size(); }
    if ((long)total > T.nz_declared) return ERR_NZMAX;       // CH.c:495-499
    const size_t vw = T.complex_field ? 2 : 1;
    T.ti.resize(total); T.tj.resize(total); T.tx.resize(total * vw);
    size_t at = 0;
    for (auto &c : chunks) {
        if (c.ti.empty()) continue;
        memcpy(T.ti.data() + at, c.ti.data(), c.ti.size() * sizeof(int64_t));
        memcpy(T.tj.data() + at, c.tj.data(), c.tj.size() * sizeof(int64_t));
        memcpy(T.tx.data() + at * vw, c.tx.data(), c.tx.size() * sizeof(double));
        at += c.ti.size();
    }
    return OK;
}

// ---- binary triplet container ----------------------------------------------------------------------------------------
// magic[8] | int64 flags (bit 0 symmetric, bit 1 complex) | int64 nrow | int64 ncol | int64 nz | Ti[nz] | Tj[nz] | Tx[nz]
// (complex: Tx holds nz (re, im) pairs)
inline bool is_binary(const char *buf, size_t len) { return len >= 8 && memcmp(buf, BIN_MAGIC, 8) == 0; }

inline int parse_binary(const char *buf, size_t len, Triplets &T) {
    if (len < 40 || !is_binary(buf, len)) return ERR_FORMAT;
    int64_t h[4];
    memcpy(h, buf + 8, sizeof(h));
    T.symmetric = (h[0] & 1) != 0;
    T.complex_field = (h[0] & 2) != 0;
    const size_t vw = T.complex_field ? 2 : 1;
    if (h[1] < 0 || h[2] < 0 || h[3] < 0 || (size_t)h[3] > (len - 40) / (16 + 8 * vw)) return ERR_FORMAT;
    T.nrow = (long)h[1]; T.ncol = (long)h[2]; T.nz_declared = (long)h[3];
    const size_t nz = (size_t)h[3];
    T.ti.resize(nz); T.tj.resize(nz); T.tx.resize(nz * vw);
    if (nz > 0) {
        memcpy(T.ti.data(), buf + 40, nz * 8);
        memcpy(T.tj.data(), buf + 40 + nz * 8, nz * 8);
        memcpy(T.tx.data(), buf + 40 + nz * 16, nz * 8 * vw);
    }
    return OK;
}

inline int write_binary(const char *path, bool symmetric, int64_t nrow, int64_t ncol, int64_t nz, const int64_t *ti,
                        const int64_t *tj, const double *tx, bool complex_field = false) {
    FILE *f = fopen(path, "wb");
    if (!f) return ERR_IO;
    const int64_t h[4] = {(symmetric ? 1 : 0) | (complex_field ? 2 : 0), nrow, ncol, nz};
    const size_t nv = (size_t)nz * (complex_field ? 2 : 1);
    bool ok = fwrite(BIN_MAGIC, 1, 8, f) == 8 && fwrite(h, 8, 4, f) == 4;
    ok = ok && fwrite(ti, 8, (size_t)nz, f) == (size_t)nz && fwrite(tj, 8, (size_t)nz, f) == (size_t)nz &&
         fwrite(tx, 8, nv, f) == nv;
    ok = (fclose(f) == 0) && ok;
    return ok ? OK : ERR_IO;
}

// whole file into memory (the FILE* is the one the caller opened, like matrix_info->file)
inline bool slurp(FILE *f, std::vector<char> &buf) {
    if (fseek(f, 0, SEEK_END) == 0) {
        const long sz = ftell(f);
        if (sz >= 0 && fseek(f, 0, SEEK_SET) == 0) {
            buf.resize((size_t)sz);
            return sz == 0 || fread(buf.data(), 1, (size_t)sz, f) == (size_t)sz;
        }
    }
    buf.clear();                                // not seekable (pipe): grow
    char tmp[1 << 16];
    size_t got;
    while ((got = fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + got);
    return !ferror(f);
}

}  // namespace sfio
