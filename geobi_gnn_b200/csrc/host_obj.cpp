// Host-side (CPU, C++) Wavefront .obj reader / writer for the file-to-file path: what the reference does
// through OpenMesh's om.read_trimesh / om.write_mesh (/root/reference/code/dataset.py:134-135, test_dual.py:29,73), for exactly
// the records the path uses - `v x y z` and `f a b c ...` (fan-triangulated, `a/b/c` tokens, negative = relative indices).
// The line rules are those of the Python reader it replaces (geobi_gnn_b200/meshio.py:_read_obj_py, kept as the cross-check):
// a record counts only if the line STARTS with "v " / "f ", tokens are separated by blanks, numbers parse as Python's
// float() / int() do for the forms a mesh file holds.  Threaded: the buffer is cut at line boundaries, one pass counts,
// one pass parses into the caller's arrays.  Built into libgeobi_host.so with g++ (no CUDA): input preparation, not the hot path.
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <charconv>
#include <string>
#include <thread>
#include <vector>

namespace {

inline bool is_eol(char c) { return c == '\n' || c == '\r'; }
inline bool is_blank(char c) { return c == ' ' || c == '\t' || c == '\v' || c == '\f'; }

struct Chunk {
    const char* begin;
    const char* end;
    int64_t n_vertices = 0, n_triangles = 0;        // counted in pass 1
    int64_t error_offset = -1;                      // byte offset of the first malformed record (pass 2)
};

// cut [buf, buf+n) into at most `parts` pieces, each starting at the beginning of a line
std::vector<Chunk> cut(const char* buf, int64_t n, int parts) {
    std::vector<Chunk> out;
    const char* p = buf;
    const char* end = buf + n;
    for (int i = 0; i < parts && p < end; ++i) {
        const char* q = (i + 1 == parts) ? end : std::max(p, buf + n / parts * (i + 1));
        while (q < end && q > buf && !is_eol(q[-1])) ++q;        // forward to just after an end-of-line
        if (q > p) out.push_back({p, q});
        p = q;
    }
    if (p < end) out.push_back({p, end});
    return out;
}

inline const char* next_line(const char* p, const char* end) {
    while (p < end && !is_eol(*p)) ++p;
    while (p < end && is_eol(*p)) ++p;
    return p;
}

inline const char* skip_blanks(const char* p, const char* eol) {
    while (p < eol && is_blank(*p)) ++p;
    return p;
}

inline const char* token_end(const char* p, const char* eol) {
    while (p < eol && !is_blank(*p)) ++p;
    return p;
}

void count_chunk(Chunk& c) {
    const char* p = c.begin;
    while (p < c.end) {
        const char* eol = p;
        while (eol < c.end && !is_eol(*eol)) ++eol;
        if (eol - p >= 2 && p[1] == ' ') {
            if (p[0] == 'v') {
                ++c.n_vertices;
            } else if (p[0] == 'f') {
                int64_t corners = 0;
                const char* q = skip_blanks(p + 1, eol);
                while (q < eol) {
                    ++corners;
                    q = skip_blanks(token_end(q, eol), eol);
                }
                if (corners > 2) c.n_triangles += corners - 2;
            }
        }
        p = next_line(eol, c.end);
    }
}

inline bool parse_double(const char* b, const char* e, double* out) {
    if (b < e && *b == '+') ++b;                      // float("+1.5") is fine in Python; from_chars rejects the sign
    auto r = std::from_chars(b, e, *out, std::chars_format::general);
    return r.ec == std::errc() && r.ptr == e;
}

inline bool parse_index(const char* b, const char* e, int64_t* out) {
    const char* slash = b;
    while (slash < e && *slash != '/') ++slash;       // "a/b/c" -> "a"
    if (b < slash && *b == '+') ++b;
    auto r = std::from_chars(b, slash, *out, 10);
    return r.ec == std::errc() && r.ptr == slash;
}

void parse_chunk(Chunk& c, const char* base, int64_t vertex_base, double* points, int64_t* faces) {
    const char* p = c.begin;
    int64_t n_seen = vertex_base;                     // vertices defined before the current line (relative indices count from here)
    std::vector<int64_t> idx;
    while (p < c.end) {
        const char* eol = p;
        while (eol < c.end && !is_eol(*eol)) ++eol;
        if (eol - p >= 2 && p[1] == ' ' && (p[0] == 'v' || p[0] == 'f')) {
            const char* q = skip_blanks(p + 1, eol);
            if (p[0] == 'v') {
                double xyz[3];
                for (int k = 0; k < 3; ++k) {
                    const char* te = token_end(q, eol);
                    if (q == te || !parse_double(q, te, &xyz[k])) {
                        if (c.error_offset < 0) c.error_offset = p - base;
                        xyz[k] = 0.0;
                    }
                    q = skip_blanks(te, eol);
                }
                memcpy(points + 3 * n_seen, xyz, sizeof(xyz));
                ++n_seen;
            } else {
                idx.clear();
                while (q < eol) {
                    const char* te = token_end(q, eol);
                    int64_t i = 0;
                    if (!parse_index(q, te, &i)) {
                        if (c.error_offset < 0) c.error_offset = p - base;
                    }
                    idx.push_back(i > 0 ? i - 1 : n_seen + i);
                    q = skip_blanks(te, eol);
                }
                for (size_t k = 1; k + 1 < idx.size(); ++k) {        // fan triangulation
                    faces[0] = idx[0];
                    faces[1] = idx[k];
                    faces[2] = idx[k + 1];
                    faces += 3;
                }
            }
        }
        p = next_line(eol, c.end);
    }
}

template <class F>
void run_parallel(size_t n, F&& fn) {
    if (n <= 1) {
        for (size_t i = 0; i < n; ++i) fn(i);
        return;
    }
    std::vector<std::thread> th;
    for (size_t i = 0; i < n; ++i) th.emplace_back([&, i] { fn(i); });
    for (auto& t : th) t.join();
}

inline char* put_int(char* p, int64_t v) {
    auto r = std::to_chars(p, p + 24, v);
    return r.ptr;
}

}  // namespace

extern "C" {

// Pass 1.  counts[0] = vertices, counts[1] = triangles after fan triangulation.  Returns 0.
__attribute__((visibility("default"))) int geobi_host_obj_count(const char* buf, int64_t n_bytes, int n_threads, int64_t* counts) {
    auto chunks = cut(buf, n_bytes, std::max(1, n_threads));
    run_parallel(chunks.size(), [&](size_t i) { count_chunk(chunks[i]); });
    counts[0] = counts[1] = 0;
    for (auto& c : chunks) {
        counts[0] += c.n_vertices;
        counts[1] += c.n_triangles;
    }
    return 0;
}

// Pass 2 into points [V,3] float64 and faces [T,3] int64 (sizes from geobi_host_obj_count with the SAME n_threads).
// Returns -1 when every record parsed, else the byte offset of the first malformed `v` / `f` line.
__attribute__((visibility("default"))) int64_t geobi_host_obj_parse(const char* buf, int64_t n_bytes, int n_threads, double* points,
                                                                   int64_t* faces) {
    auto chunks = cut(buf, n_bytes, std::max(1, n_threads));
    run_parallel(chunks.size(), [&](size_t i) { count_chunk(chunks[i]); });
    std::vector<int64_t> vbase(chunks.size() + 1, 0), tbase(chunks.size() + 1, 0);
    for (size_t i = 0; i < chunks.size(); ++i) {
        vbase[i + 1] = vbase[i] + chunks[i].n_vertices;
        tbase[i + 1] = tbase[i] + chunks[i].n_triangles;
    }
    run_parallel(chunks.size(), [&](size_t i) { parse_chunk(chunks[i], buf, vbase[i], points, faces + 3 * tbase[i]); });
    int64_t bad = -1;
    for (auto& c : chunks)
        if (c.error_offset >= 0 && (bad < 0 || c.error_offset < bad)) bad = c.error_offset;
    return bad;
}

// `# V vertices, F faces`, then `v %.6g %.6g %.6g` per vertex and 1-based `f a b c` per face (meshio.write_obj's format; %.6g is
// OpenMesh's default stream precision).  Returns 0, or -1 when the file cannot be written.
__attribute__((visibility("default"))) int geobi_host_obj_write(const char* path, const double* points, int64_t n_vertices, const int64_t* faces,
                                                               int64_t n_faces, int n_threads) {
    const int T = std::max(1, n_threads);
    std::vector<std::string> vtext(T), ftext(T);
    run_parallel(T, [&](size_t t) {
        int64_t b = n_vertices * t / T, e = n_vertices * (t + 1) / T;
        std::string& s = vtext[t];
        s.reserve((e - b) * 40);
        char line[128];
        for (int64_t i = b; i < e; ++i) {
            int n = snprintf(line, sizeof(line), "v %.6g %.6g %.6g\n", points[3 * i], points[3 * i + 1], points[3 * i + 2]);
            s.append(line, n);
        }
        b = n_faces * t / T, e = n_faces * (t + 1) / T;
        std::string& f = ftext[t];
        f.reserve((e - b) * 30);
        for (int64_t i = b; i < e; ++i) {
            char* p = line;
            *p++ = 'f';
            for (int k = 0; k < 3; ++k) {
                *p++ = ' ';
                p = put_int(p, faces[3 * i + k] + 1);
            }
            *p++ = '\n';
            f.append(line, p - line);
        }
    });
    FILE* fp = fopen(path, "wb");
    if (!fp) return -1;
    bool ok = fprintf(fp, "# %lld vertices, %lld faces\n", (long long)n_vertices, (long long)n_faces) > 0;
    for (auto& s : vtext) ok = ok && (s.empty() || fwrite(s.data(), 1, s.size(), fp) == s.size());
    for (auto& s : ftext) ok = ok && (s.empty() || fwrite(s.data(), 1, s.size(), fp) == s.size());
    ok = (fclose(fp) == 0) && ok;
    return ok ? 0 : -1;
}

}  // extern "C"
