// Host-side file formats of the hot path (C ABI in include/rcm_b200.h):
//   * repwvl lookup tables - NetCDF-4/HDF5 subset reader + flat .rcmtab reader; replaces the
//     netCDF-cxx4 calls of read_tau (reference repwvl_thermal.cpp:113-176, :208)
//   * 21-level atmosphere files (reference main.cpp:396-430)
//   * line-by-line ASCII matrices, single pass, result-identical to ASCII_file2xy2D
//     (reference lbl.arts/ascii.cpp:1631-1691 with :225-296, :1296-1337, :1391-1422)
#include <cerrno>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "rcm_internal.h"

namespace {

bool slurp(const char* path, std::vector<unsigned char>& buf) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    buf.resize(n > 0 ? (size_t)n : 0);
    size_t got = buf.empty() ? 0 : std::fread(buf.data(), 1, buf.size(), f);
    std::fclose(f);
    return got == buf.size();
}

// ---------------------------------------------------------------------------------------------
// HDF5 subset: superblock v0 with 8-byte offsets, version-2 object headers, root group links in
// a fractal heap, contiguous little-endian f64 datasets.
// ---------------------------------------------------------------------------------------------
struct H5 {
    const std::vector<unsigned char>& b;
    explicit H5(const std::vector<unsigned char>& buf) : b(buf) {}

    bool in(uint64_t off, uint64_t n) const { return off <= b.size() && n <= b.size() - off; }
    // every read is bounds-checked (a read outside the file yields 0): malformed files end in RCM_ERR_FORMAT
    unsigned at(uint64_t off) const { return off < b.size() ? b[off] : 0u; }
    uint64_t u(uint64_t off, int n) const {
        uint64_t v = 0;
        if (!in(off, (uint64_t)n)) return 0;
        for (int i = n - 1; i >= 0; --i) v = (v << 8) | at(off + i);
        return v;
    }
    bool sig(uint64_t off, const char* s) const { return in(off, 4) && std::memcmp(&b[off], s, 4) == 0; }

    struct Dataset {
        std::vector<uint64_t> shape;
        bool f64le = false;
        uint64_t addr = UINT64_MAX, size = 0;
    };

    // One link message (version 1); returns false when `p` does not start one.
    bool link(uint64_t& p, std::string& name, uint64_t& target) const {
        if (!in(p, 4) || at(p) != 1) return false;
        unsigned flags = at(p + 1);
        uint64_t q = p + 2;
        unsigned type = 0;
        if (flags & 0x08) type = at(q++);
        if (flags & 0x04) q += 8;
        if (flags & 0x10) q += 1;
        int w = 1 << (flags & 3);
        if (!in(q, w)) return false;
        uint64_t len = u(q, w);
        q += w;
        if (len == 0 || len > b.size() || !in(q, len + 8)) return false;
        name.assign((const char*)&b[q], len);
        q += len;
        target = UINT64_MAX;
        if (type == 0) {
            target = u(q, 8);
            q += 8;
        } else {
            return false;  // soft/external links do not occur in these tables
        }
        p = q;
        return true;
    }

    void messages(uint64_t p, uint64_t end, bool tracked, Dataset& d, std::map<std::string, uint64_t>& links,
                  int depth) const {
        while (p + 4 <= end && in(p, 4)) {
            unsigned type = at(p);
            uint64_t size = u(p + 1, 2);
            uint64_t body = p + 4 + (tracked ? 2 : 0);
            if (!in(body, size)) return;
            if (type == 0x01) {  // dataspace
                unsigned ver = at(body), rank = at(body + 1);
                uint64_t q = body + (ver == 2 ? 4 : 8);
                d.shape.clear();
                for (unsigned i = 0; i < rank && in(q, 8); ++i, q += 8) d.shape.push_back(u(q, 8));
            } else if (type == 0x03) {  // datatype: class 1 = IEEE float, bit0 of the class bits = big endian
                unsigned cls = at(body) & 0x0F, bits0 = at(body + 1);
                d.f64le = (cls == 1 && u(body + 4, 4) == 8 && (bits0 & 1) == 0);
            } else if (type == 0x08) {  // layout: version 3, class 1 = contiguous
                if (at(body) == 3 && at(body + 1) == 1) {
                    d.addr = u(body + 2, 8);
                    d.size = u(body + 10, 8);
                }
            } else if (type == 0x06) {  // compact link
                uint64_t q = body, tgt;
                std::string nm;
                if (link(q, nm, tgt)) links[nm] = tgt;
            } else if (type == 0x10 && depth < 16) {  // continuation block
                uint64_t off = u(body, 8), len = u(body + 8, 8);
                if (sig(off, "OCHK") && in(off, len) && len >= 8) messages(off + 4, off + len - 4, tracked, d, links, depth + 1);
            }
            p = body + size;
        }
    }

    bool object(uint64_t addr, Dataset& d, std::map<std::string, uint64_t>& links) const {
        if (!sig(addr, "OHDR") || at(addr + 4) != 2) return false;
        unsigned flags = at(addr + 5);
        uint64_t p = addr + 6;
        if (flags & 0x20) p += 16;
        if (flags & 0x10) p += 4;
        int w = 1 << (flags & 3);
        uint64_t chunk0 = u(p, w);
        p += w;
        if (!in(p, chunk0)) return false;
        messages(p, p + chunk0, (flags & 0x04) != 0, d, links, 0);
        return true;
    }

    // Dense link storage: decode the link messages stored back to back in the managed
    // direct blocks of every fractal heap.
    void heap_links(std::map<std::string, uint64_t>& links) const {
        for (uint64_t h = 0; h + 4 <= b.size(); ++h) {
            if (!sig(h, "FRHP") || at(h + 4) != 0) continue;
            unsigned hflags = at(h + 9);
            // fixed part of the heap header up to "maximum heap size" (bits)
            uint64_t maxbits_off = h + 4 + 1 + 2 + 2 + 1 + 4 + 8 * 12 + 2 + 8 + 8;
            if (!in(maxbits_off, 2)) continue;
            uint64_t off_bytes = (u(maxbits_off, 2) + 7) / 8;
            for (uint64_t d = 0; d + 4 <= b.size(); ++d) {
                if (!sig(d, "FHDB") || at(d + 4) != 0 || u(d + 5, 8) != h) continue;
                uint64_t p = d + 5 + 8 + off_bytes + ((hflags & 0x02) ? 4 : 0);
                std::string nm;
                uint64_t tgt;
                while (link(p, nm, tgt)) links[nm] = tgt;
            }
        }
    }
};

int load_nc4(const std::vector<unsigned char>& buf, rcm_table& t) {
    static const unsigned char magic[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
    if (buf.size() < 96 || std::memcmp(buf.data(), magic, 8) != 0) return RCM_ERR_FORMAT;
    if (buf[8] != 0 || buf[13] != 8 || buf[14] != 8) return RCM_ERR_FORMAT;  // superblock v0, 8-byte offsets/lengths
    H5 h5(buf);
    std::map<std::string, uint64_t> links;
    h5.heap_links(links);
    {  // small files keep the links compact in the root object header (root symbol-table entry at 24+32)
        H5::Dataset dummy;
        uint64_t root = h5.u(24 + 32 + 8, 8);
        h5.object(root, dummy, links);
    }
    auto fetch = [&](const char* name, std::vector<double>& out, std::vector<uint64_t>& shape) -> bool {
        auto it = links.find(name);
        if (it == links.end()) return false;
        H5::Dataset d;
        std::map<std::string, uint64_t> ignore;
        if (!h5.object(it->second, d, ignore) || !d.f64le || d.addr == UINT64_MAX) return false;
        uint64_t n = 1;
        for (uint64_t s : d.shape) n *= s;
        if (n == 0 || d.size != 8 * n || !h5.in(d.addr, d.size)) return false;
        out.resize(n);
        std::memcpy(out.data(), &buf[d.addr], d.size);  // little-endian host assumed (x86-64)
        shape = d.shape;
        return true;
    };
    std::vector<uint64_t> sx, s1;
    if (!fetch("xsec", t.xsec, sx) || sx.size() != 4) return RCM_ERR_FORMAT;
    t.n_tpert = (int)sx[0];
    t.n_species = (int)sx[1];
    t.n_wvl = (int)sx[2];
    t.n_p = (int)sx[3];
    if (!fetch("ChosenWvls", t.wvl, s1) || (int)t.wvl.size() != t.n_wvl) return RCM_ERR_FORMAT;
    if (!fetch("ChosenWeights", t.weight, s1) || (int)t.weight.size() != t.n_wvl) return RCM_ERR_FORMAT;
    if (!fetch("p_grid", t.p_grid, s1) || (int)t.p_grid.size() != t.n_p) return RCM_ERR_FORMAT;
    if (!fetch("t_ref", t.t_ref, s1) || (int)t.t_ref.size() != t.n_p) return RCM_ERR_FORMAT;
    if (!fetch("t_pert", t.t_pert, s1) || (int)t.t_pert.size() != t.n_tpert) return RCM_ERR_FORMAT;
    if (!fetch("vmrs_ref", t.vmrs_ref, s1)) t.vmrs_ref.clear();  // read but never used by the reference (:171)
    return RCM_OK;
}

int load_rcmtab(const std::vector<unsigned char>& buf, rcm_table& t) {
    if (buf.size() < 40 || std::memcmp(buf.data(), "RCMTAB01", 8) != 0) return RCM_ERR_FORMAT;
    uint64_t d[4];
    std::memcpy(d, &buf[8], 32);
    for (int i = 0; i < 4; ++i)
        if (d[i] == 0 || d[i] > (1u << 24)) return RCM_ERR_FORMAT;
    t.n_tpert = (int)d[0];
    t.n_species = (int)d[1];
    t.n_wvl = (int)d[2];
    t.n_p = (int)d[3];
    size_t off = 40;
    auto take = [&](std::vector<double>& v, size_t n) -> bool {
        if (buf.size() - off < 8 * n) return false;
        v.resize(n);
        std::memcpy(v.data(), &buf[off], 8 * n);
        off += 8 * n;
        return true;
    };
    bool ok = take(t.xsec, (size_t)d[0] * d[1] * d[2] * d[3]) && take(t.wvl, d[2]) && take(t.weight, d[2]) &&
              take(t.p_grid, d[3]) && take(t.t_ref, d[3]) && take(t.t_pert, d[0]) && take(t.vmrs_ref, d[1] * d[3]);
    return ok ? RCM_OK : RCM_ERR_FORMAT;
}

}  // namespace

extern "C" {

int rcm_table_load(const char* path, rcm_table** out) {
    if (!path || !out) return RCM_ERR_ARG;
    *out = nullptr;
    std::vector<unsigned char> buf;
    if (!slurp(path, buf)) return RCM_ERR_IO;
    rcm_table* t = new (std::nothrow) rcm_table();
    if (!t) return RCM_ERR_NOMEM;
    int st = (buf.size() >= 8 && std::memcmp(buf.data(), "RCMTAB01", 8) == 0) ? load_rcmtab(buf, *t) : load_nc4(buf, *t);
    if (st != RCM_OK) {
        delete t;
        return st;
    }
    *out = t;
    return RCM_OK;
}

void rcm_table_free(rcm_table* t) { delete t; }

int rcm_table_dims(const rcm_table* t, int* dims4) {
    if (!t || !dims4) return RCM_ERR_ARG;
    dims4[0] = t->n_tpert;
    dims4[1] = t->n_species;
    dims4[2] = t->n_wvl;
    dims4[3] = t->n_p;
    return RCM_OK;
}

const double* rcm_table_array(const rcm_table* t, int which) {
    if (!t) return nullptr;
    switch (which) {
        case 0: return t->xsec.data();
        case 1: return t->wvl.data();
        case 2: return t->weight.data();
        case 3: return t->p_grid.data();
        case 4: return t->t_ref.data();
        case 5: return t->t_pert.data();
        case 6: return t->vmrs_ref.empty() ? nullptr : t->vmrs_ref.data();
        default: return nullptr;
    }
}

// main.cpp:396-430: four header lines are dropped unconditionally, then every line is split
// with operator>> semantics; values go to per-column arrays.
int rcm_read_atm(const char* path, int max_rows, double* cols_out, int* nrows_out, int* ncols_out) {
    if (!path || !cols_out || !nrows_out || max_rows <= 0) return RCM_ERR_ARG;
    FILE* f = std::fopen(path, "r");
    if (!f) return RCM_ERR_IO;
    std::vector<char> line(1 << 16);
    for (int i = 0; i < 4; ++i)
        if (!std::fgets(line.data(), (int)line.size(), f)) break;
    int rows = 0, maxc = 0;
    while (std::fgets(line.data(), (int)line.size(), f)) {
        char* p = line.data();
        int c = 0;
        while (c < 9) {
            char* e = nullptr;
            double v = std::strtod(p, &e);
            if (e == p) break;
            if (rows < max_rows) cols_out[(size_t)c * max_rows + rows] = v;
            p = e;
            ++c;
        }
        if (c == 0) continue;
        if (rows >= max_rows) {
            std::fclose(f);
            return RCM_ERR_ARG;
        }
        if (c > maxc) maxc = c;
        ++rows;
    }
    std::fclose(f);
    *nrows_out = rows;
    if (ncols_out) *ncols_out = maxc;
    return rows > 0 ? RCM_OK : RCM_ERR_FORMAT;
}

// Single-pass equivalent of ASCII_file2xy2D.  Tokenisation as the reference: delimiters are
// exactly ' ', '\t', '\n' (ascii.cpp:257); a line whose first token starts with '%' or '#' is
// skipped, a later token starting with one ends the row (ascii.cpp:163-164, :259-276); every
// cell goes through strtod (:1404-1414); a non-rectangular or empty matrix returns -5 (:1653).
int rcm_ascii_file2xy2D(const char* filename, int* nx, int* ny, double** x, double** y) {
    if (!filename || !nx || !ny || !x || !y) return RCM_ERR_ARG;
    *nx = *ny = 0;
    *x = *y = nullptr;
    std::vector<unsigned char> buf;
    {
        FILE* f = std::fopen(filename, "r");
        if (!f) return -1;  // ASCIIFILE_NOT_FOUND
        std::fclose(f);
    }
    if (!slurp(filename, buf)) return -1;
    buf.push_back('\n');
    std::vector<double> vals;
    vals.reserve(buf.size() / 8);
    long rows = 0;
    int mincol = INT_MAX, maxcol = 0;
    auto delim = [](unsigned char c) { return c == ' ' || c == '\t' || c == '\n'; };
    size_t p = 0, n = buf.size();
    while (p < n) {
        size_t eol = p;
        while (buf[eol] != '\n') ++eol;
        buf[eol] = 0;
        int cols = 0;
        size_t q = p;
        while (q < eol) {
            while (q < eol && delim(buf[q])) ++q;
            if (q >= eol) break;
            size_t s = q;
            while (q < eol && !delim(buf[q])) ++q;
            if (buf[s] == '%' || buf[s] == '#') break;
            unsigned char keep = buf[q];
            buf[q] = 0;
            vals.push_back(std::strtod((const char*)&buf[s], nullptr));
            buf[q] = keep;
            ++cols;
        }
        if (cols > 0) {
            ++rows;
            if (cols < mincol) mincol = cols;
            if (cols > maxcol) maxcol = cols;
        }
        p = eol + 1;
    }
    if (mincol != maxcol) return -5;  // NOT_A_RECTANGULAR_MATRIX (also what an empty file yields)
    const long nc = maxcol - 1;
    double* xx = (double*)std::calloc((size_t)rows, sizeof(double));
    double* yy = (double*)std::calloc((size_t)rows * (nc > 0 ? nc : 1), sizeof(double));
    if (!xx || !yy) {
        std::free(xx);
        std::free(yy);
        return -2;  // ASCII_NO_MEMORY
    }
    for (long r = 0; r < rows; ++r) {
        xx[r] = vals[(size_t)r * maxcol];
        for (long c = 0; c < nc; ++c) yy[(size_t)r * nc + c] = vals[(size_t)r * maxcol + 1 + c];
    }
    *nx = (int)rows;
    *ny = (int)nc;
    *x = xx;
    *y = yy;
    return 0;
}

void rcm_free(void* p) { std::free(p); }

// Batched output_conv (main.cpp:102-114): see include/rcm_b200.h.
int rcm_write_profiles(const char* path, int append, int header, int ncol, const double* plevel_hPa,
                       const double* Tlayer, const float* time_h, int column_ids) {
    if (!path || ncol < 0 || !plevel_hPa || (ncol > 0 && (!Tlayer || !time_h))) return RCM_ERR_ARG;
    double player[RCM_NLAYER], conv[RCM_NLAYER];
    for (int l = 0; l < RCM_NLAYER; ++l) {
        player[l] = (plevel_hPa[l] + plevel_hPa[l + 1]) / 2.0;   // main.cpp:472
        conv[l] = std::pow(1000.0 / player[l], 2.0 / 7.0);       // main.cpp:474
    }
    FILE* f = std::fopen(path, append ? "a" : "w");
    if (!f) return RCM_ERR_IO;
    std::string out;
    out.reserve(1 << 20);
    char row[1600];  // "%f" of a diverged temperature (1e300) is over 300 characters: four of them still fit
    bool ok = true;
    if (header) out += column_ids ? "column,layer,player,Tlayer,theta,time\n" : "layer,player,Tlayer,theta,time\n";
    for (int c = 0; c < ncol && ok; ++c) {
        const double* T = Tlayer + (size_t)c * RCM_NLAYER;
        for (int l = 0; l < RCM_NLAYER; ++l) {
            const double theta = T[l] * conv[l];  // t_to_theta, main.cpp:125
            int n = 0;
            if (column_ids) n = std::snprintf(row, sizeof(row), "%d,", c);
            const int m = std::snprintf(row + n, sizeof(row) - n, "%d,%f,%f,%f,%f\n", l, player[l], T[l], theta, (double)time_h[c]);
            if (n < 0 || m < 0 || (size_t)(n + m) >= sizeof(row)) {  // cannot happen with finite doubles; never read past `row`
                ok = false;
                break;
            }
            out.append(row, (size_t)(n + m));
        }
        if (out.size() > (1u << 20) - 4096) {
            ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
            out.clear();
        }
    }
    if (ok && !out.empty()) ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    ok = (std::fclose(f) == 0) && ok;
    return ok ? RCM_OK : RCM_ERR_IO;
}

}  // extern "C"
