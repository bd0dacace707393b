// Trajectory ingest: LAMMPS text dump -> (T, N, ndim) doubles in a caller-owned host buffer (pinned, so that
// MeshPipeline.run_from_host can stream it to the GPUs).  Host-only code: the reference parses the dump line by line
// in Python (read_lammps_trajectory, dynaphopy/interface/iofile/trajectory_parsers.py:135-352: mmap.find for the
// "TIMESTEP" / "NUMBER OF ATOMS" / "BOX BOUNDS" / "ITEM: ATOMS" markers, then `readline().split()[0:ndim]` and
// float() for every atom of every frame); here one scan finds the frames and a team of threads converts them with
// strtod (correctly rounded, like Python's float()).  Semantics kept: frames are counted from 1 at every TIMESTEP
// marker, the first `ndim` columns of every atom line are taken whatever the column labels say, the cell comes from
// the FIRST "BOX BOUNDS" block, the number of atoms from the FIRST "NUMBER OF ATOMS" block.
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "dp_common.cuh"

namespace dp {

struct MappedFile {
    const char* p = nullptr;
    size_t n = 0, mapped = 0;
    int fd = -1;
    ~MappedFile() {
        if (p && mapped) munmap((void*)p, mapped);
        if (fd >= 0) close(fd);
    }
    int open_ro(const char* path) {
        fd = ::open(path, O_RDONLY);
        DP_REQUIRE(fd >= 0, DP_ERR_INVALID, "Trajectory file does not exist! (%s: %s)", path, strerror(errno));
        struct stat st;
        DP_REQUIRE(fstat(fd, &st) == 0, DP_ERR_INVALID, "cannot stat %s", path);
        n = (size_t)st.st_size;
        if (n == 0) return DP_OK;
        // strtod / strtol need a terminator: reserve n + 1 bytes (rounded to pages) of anonymous zero memory and map the
        // file over its start, so the byte after the last one of the file is always a readable NUL -- also when the
        // file size is a multiple of the page size and the file ends inside a number (no trailing newline)
        const size_t page = (size_t)sysconf(_SC_PAGESIZE);
        mapped = (n + 1 + page - 1) / page * page;
        void* m = mmap(nullptr, mapped, PROT_READ, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        DP_REQUIRE(m != MAP_FAILED, DP_ERR_INVALID, "cannot reserve %zu bytes for %s: %s", mapped, path, strerror(errno));
        p = (const char*)m;
        void* f = mmap(m, n, PROT_READ, MAP_PRIVATE | MAP_FIXED, fd, 0);
        DP_REQUIRE(f == m, DP_ERR_INVALID, "cannot map %s: %s", path, strerror(errno));
        madvise(m, n, MADV_SEQUENTIAL);
        return DP_OK;
    }
};

static const char* find_token(const char* b, const char* e, const char* tok) {
    return (const char*)memmem(b, (size_t)(e - b), tok, strlen(tok));
}
static const char* next_line(const char* b, const char* e) {       // first character after the next '\n' (or e)
    const char* nl = (const char*)memchr(b, '\n', (size_t)(e - b));
    return nl ? nl + 1 : e;
}

struct LammpsIndex {
    std::vector<size_t> data_off;     // offset of the first atom line of every frame
    std::vector<double> timestep;     // the number on the line after the TIMESTEP marker
    std::string labels;               // the "ITEM: ATOMS ..." line of the last frame scanned (reference: lammps_labels)
};

static int scan_lammps(const MappedFile& f, LammpsIndex* ix) {
    const char* b = f.p;
    const char* e = f.p + f.n;
    const char* cur = b;
    while (cur < e) {
        const char* ts = find_token(cur, e, "TIMESTEP");
        if (!ts) break;
        const char* l = next_line(ts, e);
        DP_REQUIRE(l < e, DP_ERR_INVALID, "LAMMPS dump: truncated after a TIMESTEP marker");
        char* endp = nullptr;
        const double t = strtod(l, &endp);
        DP_REQUIRE(endp != l, DP_ERR_INVALID, "LAMMPS dump: no time step value in frame %zu", ix->timestep.size() + 1);
        const char* at = find_token(l, e, "ITEM: ATOMS");
        if (!at) break;                                   // a TIMESTEP header without atoms: incomplete last frame
        const char* data = next_line(at, e);
        ix->labels.assign(at, (size_t)(data - at));
        ix->timestep.push_back(t);
        ix->data_off.push_back((size_t)(data - b));
        cur = data;
    }
    return DP_OK;
}

// parse `natoms` lines of `ndim` leading numbers starting at p; returns the first unparsable atom (or -1)
static long parse_frame(const char* p, const char* e, int natoms, int ndim, double* out) {
    for (int i = 0; i < natoms; i++) {
        if (p >= e) return i;
        const char* le = (const char*)memchr(p, '\n', (size_t)(e - p));
        if (!le) le = e;
        const char* q = p;
        for (int k = 0; k < ndim; k++) {
            char* endp = nullptr;
            // strtod stops at the newline at the latest only if the line holds a number: guard against running into
            // the next line by checking the end position
            while (q < le && (*q == ' ' || *q == '\t' || *q == '\r')) q++;
            if (q >= le) return i;
            const double v = strtod(q, &endp);
            if (endp == q || endp > le) return i;
            out[(size_t)i * ndim + k] = v;
            q = endp;
        }
        p = le < e ? le + 1 : e;
    }
    return -1;
}

// VASP XDATCAR (read_VASP_XDATCAR, trajectory_parsers.py:355-483): two header lines, three cell lines, one species
// line, the atom counts; then "Direct configuration= k" followed by N lines of fractional coordinates.
static int scan_xdatcar(const MappedFile& f, LammpsIndex* ix, int* n_atoms, double* cell) {
    const char* b = f.p;
    const char* e = f.p + f.n;
    const char* l = b;
    for (int i = 0; i < 2; i++) l = next_line(l, e);
    for (int r = 0; r < 3; r++) {
        const char* le = next_line(l, e);
        const char* q = l;
        for (int c = 0; c < 3; c++) {
            char* endp = nullptr;
            cell[r * 3 + c] = strtod(q, &endp);
            DP_REQUIRE(endp != q && endp <= le, DP_ERR_INVALID, "XDATCAR: bad cell line %d", r + 1);
            q = endp;
        }
        l = le;
    }
    l = next_line(l, e);                                   // species names
    {
        const char* le = next_line(l, e);
        long total = 0;
        const char* q = l;
        while (q < le) {
            char* endp = nullptr;
            const long v = strtol(q, &endp, 10);
            if (endp == q || endp > le) break;
            total += v;
            q = endp;
        }
        DP_REQUIRE(total > 0, DP_ERR_INVALID, "XDATCAR: no atom counts in the header");
        *n_atoms = (int)total;
        l = le;
    }
    const char* cur = l;
    while (cur < e) {
        const char* m = find_token(cur, e, "Direct configuration");
        if (!m) break;
        const char* le = next_line(m, e);
        const char* eq = (const char*)memchr(m, '=', (size_t)(le - m));
        DP_REQUIRE(eq != nullptr, DP_ERR_INVALID, "XDATCAR: no '=' in the configuration line of frame %zu", ix->timestep.size() + 1);
        ix->timestep.push_back(strtod(eq + 1, nullptr));
        ix->data_off.push_back((size_t)(le - b));
        cur = le;
    }
    return DP_OK;
}

// frames [first, first + n) of an index -> out, all threads; error text names the first bad step (1-based)
static int convert_frames(const MappedFile& f, const LammpsIndex& ix, int64_t first_frame, int64_t n_frames, int n_atoms,
                          int ndim, double* out, int n_threads) {
    const int nt = (int)imax<int64_t>(1, imin<int64_t>(n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency(), n_frames));
    std::vector<long> bad_atom((size_t)nt, -1);
    std::vector<int64_t> bad_frame((size_t)nt, -1);
    auto work = [&](int w) {
        for (int64_t i = w; i < n_frames; i += nt) {
            const size_t off = ix.data_off[(size_t)(first_frame + i)];
            const long r = parse_frame(f.p + off, f.p + f.n, n_atoms, ndim, out + (size_t)i * n_atoms * ndim);
            if (r >= 0 && bad_frame[(size_t)w] < 0) { bad_frame[(size_t)w] = first_frame + i; bad_atom[(size_t)w] = r; }
        }
    };
    std::vector<std::thread> team;
    for (int w = 1; w < nt; w++) team.emplace_back(work, w);
    work(0);
    for (auto& t : team) t.join();
    int64_t first_bad = -1;
    long atom = -1;
    for (int w = 0; w < nt; w++)
        if (bad_frame[(size_t)w] >= 0 && (first_bad < 0 || bad_frame[(size_t)w] < first_bad)) { first_bad = bad_frame[(size_t)w]; atom = bad_atom[(size_t)w]; }
    // the reference stops with "Error reading step N" (trajectory_parsers.py:312-315, 452-454)
    DP_REQUIRE(first_bad < 0, DP_ERR_INVALID, "Error reading step %lld (atom line %ld)", (long long)(first_bad + 1), atom + 1);
    return DP_OK;
}

}  // namespace dp

extern "C" {

int dp_xdatcar_scan(const char* path, int64_t* n_frames, int* n_atoms, double* cell) {
    DP_REQUIRE(path && n_frames && n_atoms && cell, DP_ERR_INVALID, "dp_xdatcar_scan: null pointer");
    dp::MappedFile f;
    int rc = f.open_ro(path);
    if (rc) return rc;
    DP_REQUIRE(f.n > 0, DP_ERR_INVALID, "XDATCAR: empty file");
    dp::LammpsIndex ix;
    rc = dp::scan_xdatcar(f, &ix, n_atoms, cell);
    if (rc) return rc;
    *n_frames = (int64_t)ix.timestep.size();
    return DP_OK;
}

int dp_xdatcar_read(const char* path, int64_t first_frame, int64_t n_frames, int n_atoms, double* out, double* steps,
                    int64_t n_steps, int n_threads) {
    DP_REQUIRE(path && (out || n_frames == 0), DP_ERR_INVALID, "dp_xdatcar_read: null pointer");
    DP_REQUIRE(first_frame >= 0 && n_frames >= 0 && n_atoms > 0, DP_ERR_INVALID, "dp_xdatcar_read: bad sizes");
    dp::MappedFile f;
    int rc = f.open_ro(path);
    if (rc) return rc;
    DP_REQUIRE(f.n > 0, DP_ERR_INVALID, "XDATCAR: empty file");
    dp::LammpsIndex ix;
    int na = 0;
    double cell[9];
    rc = dp::scan_xdatcar(f, &ix, &na, cell);
    if (rc) return rc;
    DP_REQUIRE(first_frame + n_frames <= (int64_t)ix.timestep.size(), DP_ERR_INVALID,
               "dp_xdatcar_read: frames [%lld, %lld) requested, the file holds %zu", (long long)first_frame,
               (long long)(first_frame + n_frames), ix.timestep.size());
    // the configuration numbers of frames [0, n_steps): the reference keeps the time of the skipped frames too (:418)
    if (steps)
        for (int64_t i = 0; i < n_steps && i < (int64_t)ix.timestep.size(); i++) steps[i] = ix.timestep[(size_t)i];
    return dp::convert_frames(f, ix, first_frame, n_frames, n_atoms, 3, out, n_threads);
}

int dp_lammps_scan(const char* path, int64_t* n_frames, int* n_atoms, double* bounds, int* bound_cols, char* labels,
                   int labels_len) {
    DP_REQUIRE(path && n_frames && n_atoms && bounds && bound_cols, DP_ERR_INVALID, "dp_lammps_scan: null pointer");
    dp::MappedFile f;
    int rc = f.open_ro(path);
    if (rc) return rc;
    const char* b = f.p;
    const char* e = f.p + f.n;
    dp::LammpsIndex ix;
    rc = dp::scan_lammps(f, &ix);
    if (rc) return rc;
    *n_frames = (int64_t)ix.timestep.size();
    *n_atoms = 0;
    *bound_cols = 0;
    for (int i = 0; i < 9; i++) bounds[i] = 0.0;
    if (f.n) {
        if (const char* na = dp::find_token(b, e, "NUMBER OF ATOMS")) {
            const char* l = dp::next_line(na, e);
            *n_atoms = (int)strtol(l, nullptr, 10);
        }
        if (const char* bb = dp::find_token(b, e, "BOX BOUNDS")) {
            const char* l = dp::next_line(bb, e);
            int cols = 3;
            for (int r = 0; r < 3 && l < e; r++) {
                const char* le = (const char*)memchr(l, '\n', (size_t)(e - l));
                if (!le) le = e;
                int c = 0;
                const char* q = l;
                while (c < 3) {
                    while (q < le && (*q == ' ' || *q == '\t' || *q == '\r')) q++;
                    if (q >= le) break;
                    char* endp = nullptr;
                    const double v = strtod(q, &endp);
                    if (endp == q || endp > le) break;
                    bounds[r * 3 + c++] = v;
                    q = endp;
                }
                DP_REQUIRE(c >= 2, DP_ERR_INVALID, "LAMMPS dump: bad BOX BOUNDS line %d", r + 1);
                if (c < cols) cols = c;
                l = le < e ? le + 1 : e;
            }
            *bound_cols = cols;
        }
    }
    if (labels && labels_len > 0) {
        const size_t n = dp::imin<size_t>(ix.labels.size(), (size_t)labels_len - 1);
        memcpy(labels, ix.labels.data(), n);
        labels[n] = 0;
    }
    return DP_OK;
}

int dp_lammps_read(const char* path, int64_t first_frame, int64_t n_frames, int n_atoms, int ndim, double* out,
                   double* timesteps, int n_threads) {
    DP_REQUIRE(path && (out || n_frames == 0), DP_ERR_INVALID, "dp_lammps_read: null pointer");
    DP_REQUIRE(first_frame >= 0 && n_frames >= 0 && n_atoms > 0 && ndim >= 1 && ndim <= 3, DP_ERR_INVALID,
               "dp_lammps_read: bad sizes");
    dp::MappedFile f;
    int rc = f.open_ro(path);
    if (rc) return rc;
    dp::LammpsIndex ix;
    rc = dp::scan_lammps(f, &ix);
    if (rc) return rc;
    DP_REQUIRE(first_frame + n_frames <= (int64_t)ix.timestep.size(), DP_ERR_INVALID,
               "dp_lammps_read: frames [%lld, %lld) requested, the dump holds %zu", (long long)first_frame,
               (long long)(first_frame + n_frames), ix.timestep.size());
    if (timesteps)
        for (int64_t i = 0; i < n_frames; i++) timesteps[i] = ix.timestep[(size_t)(first_frame + i)];
    return dp::convert_frames(f, ix, first_frame, n_frames, n_atoms, ndim, out, n_threads);
}

}  // extern "C"
