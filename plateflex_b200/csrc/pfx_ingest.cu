// pfx_ingest.cu -- the step in front of the hot path: reading a GMT ".xyz" grid dump, the
// nearest-neighbour fill of NaN cells that Grid.__init__ performs (classes.py:143-153, scipy griddata
// method='nearest') and the Gaussian smoothing behind TopoGrid.filter_water_depth (classes.py:538-546,
// skimage.filters.gaussian = scipy.ndimage.gaussian_filter with mode='nearest', truncate=4).
//
//   * the reader is a plain C++ parser over the mapped file, split over host threads at line boundaries (the
//     reference's notebooks go through pandas.read_csv);
//   * the fill is an exact two-pass nearest-valid-cell search on the GPU: per column the nearest valid row above
//     and below every cell (one scan), then per NaN cell the minimum over columns of dj^2 + di(j')^2;
//   * the Gaussian is two separable FP64 passes with edge replication.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "pfx_common.cuh"

namespace pfx {
namespace {

struct Mapped {
    const char *p = nullptr;
    size_t n = 0;
    int fd = -1;
    ~Mapped()
    {
        if (p) munmap(const_cast<char *>(p), n);
        if (fd >= 0) close(fd);
    }
    bool open_file(const char *path)
    {
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0 || st.st_size == 0) return false;
        n = (size_t)st.st_size;
        void *m = mmap(nullptr, n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) return false;
        p = static_cast<const char *>(m);
        return true;
    }
};

// end of the header: the first line when it starts with '#'
size_t header_end(const Mapped &f)
{
    if (f.n == 0 || f.p[0] != '#') return 0;
    const char *e = static_cast<const char *>(memchr(f.p, '\n', f.n));
    return e ? (size_t)(e - f.p) + 1 : f.n;
}

// parse "x y z" lines of [b, e) into z (the third field of each non-empty line); returns the number parsed or -1
long parse_chunk(const char *b, const char *e, double *z, long cap)
{
    long n = 0;
    const char *p = b;
    while (p < e) {
        while (p < e && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\n')) p++;
        if (p >= e) break;
        if (*p == '#' || *p == '>') {  // comment / GMT segment header
            while (p < e && *p != '\n') p++;
            continue;
        }
        char *q = nullptr;
        double v[3];
        for (int c = 0; c < 3; c++) {
            v[c] = strtod(p, &q);  // accepts "NaN"
            if (q == p) return -1;
            p = q;
        }
        if (n >= cap) return -1;
        z[n++] = v[2];
        while (p < e && *p != '\n') p++;
    }
    return n;
}

// ---- nearest-valid-cell fill -------------------------------------------------------------------------
// data is (nx, ny) C order (y fastest).  up/down: for every cell the distance (in rows) to the nearest valid
// cell of its column at or above / at or below it, or a large value when there is none.
constexpr int FAR = 1 << 29;

__global__ void column_scan_kernel(int nx, int ny, const double *__restrict__ d, int *__restrict__ near)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ny) return;
    int last = -FAR;
    for (int i = 0; i < nx; i++) {  // nearest valid at or above
        if (isfinite(d[(size_t)i * ny + j])) last = i;
        near[(size_t)i * ny + j] = (last == -FAR) ? FAR : i - last;
    }
    last = FAR;
    for (int i = nx - 1; i >= 0; i--) {  // combine with the nearest valid at or below (ties: the one above)
        if (isfinite(d[(size_t)i * ny + j])) last = i;
        const int up = near[(size_t)i * ny + j], dn = (last == FAR) ? FAR : last - i;
        near[(size_t)i * ny + j] = (dn < up) ? -dn : up;  // signed: negative = below
    }
}

// one warp per NaN cell: minimum over columns j' of (j - j')^2 + near(i, j')^2; ties go to the smallest
// (distance, row, column) so that the result does not depend on the schedule
__global__ void fill_kernel(int nx, int ny, const double *__restrict__ src, const int *__restrict__ near,
                            const int *__restrict__ todo, int ntodo, double *__restrict__ dst)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= ntodo) return;
    const int c = todo[w], i = c / ny, j = c - i * ny;
    long long bd = (1ll << 62);
    int bi = 0, bj = 0;
    for (int jj = lane; jj < ny; jj += 32) {
        const int s = near[(size_t)i * ny + jj];
        if (s == FAR) continue;
        const long long di = s < 0 ? -s : s, dj = jj - j;
        const long long d2 = di * di + dj * dj;
        const int ii = i - s;  // s > 0: valid cell above, s < 0: below (row i + |s|)
        if (d2 < bd || (d2 == bd && (ii < bi || (ii == bi && jj < bj)))) {
            bd = d2;
            bi = ii;
            bj = jj;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const long long od = __shfl_down_sync(0xffffffffu, bd, o);
        const int oi = __shfl_down_sync(0xffffffffu, bi, o), oj = __shfl_down_sync(0xffffffffu, bj, o);
        if (od < bd || (od == bd && (oi < bi || (oi == bi && oj < bj)))) {
            bd = od;
            bi = oi;
            bj = oj;
        }
    }
    if (lane == 0 && bd < (1ll << 62)) dst[c] = src[(size_t)bi * ny + bj];
}

// ---- separable Gaussian, edge replication (scipy.ndimage mode='nearest') ------------------------------
__global__ void gauss_axis_kernel(int nx, int ny, int axis, int radius, const double *__restrict__ w,
                                  const double *__restrict__ in, double *__restrict__ out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= ny) return;
    double acc = 0.0;
    if (axis == 0) {
        for (int t = -radius; t <= radius; t++) {
            const int ii = min(max(i + t, 0), nx - 1);
            acc += w[t + radius] * in[(size_t)ii * ny + j];
        }
    } else {
        for (int t = -radius; t <= radius; t++) {
            const int jj = min(max(j + t, 0), ny - 1);
            acc += w[t + radius] * in[(size_t)i * ny + jj];
        }
    }
    out[(size_t)i * ny + j] = acc;
}

struct DevGuard {
    int prev = -1;
    explicit DevGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DevGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

}  // namespace
}  // namespace pfx

using namespace pfx;

extern "C" {

int pfx_xyz_header(const char *path, double *hdr10)
{
    PFX_ARG(path && hdr10, "null pointer");
    Mapped f;
    if (!f.open_file(path)) {
        set_error("cannot open %s", path);
        return PFX_ERR_ARG;
    }
    const size_t he = header_end(f);
    PFX_ARG(he > 0, "the file does not start with a '# ...' header line (gmt grdinfo -C)");
    // "# name<TAB>xmin xmax ymin ymax zmin zmax dx dy nx ny": the last ten numeric fields of the line
    std::vector<double> nums;
    std::string line(f.p + 1, he - 1);
    size_t pos = 0;
    while (pos < line.size()) {
        while (pos < line.size() && (line[pos] == ' ' || line[pos] == '\t' || line[pos] == '\r' || line[pos] == '\n')) pos++;
        if (pos >= line.size()) break;
        size_t end = pos;
        while (end < line.size() && line[end] != ' ' && line[end] != '\t' && line[end] != '\r' && line[end] != '\n') end++;
        const std::string tok = line.substr(pos, end - pos);
        char *q = nullptr;
        const double v = strtod(tok.c_str(), &q);
        if (q && *q == '\0' && q != tok.c_str()) nums.push_back(v);
        else nums.clear();  // a non-numeric field (the grid name) restarts the run
        pos = end;
    }
    PFX_ARG(nums.size() >= 10, "header holds fewer than ten numeric fields");
    for (int c = 0; c < 10; c++) hdr10[c] = nums[nums.size() - 10 + c];
    return PFX_OK;
}

int pfx_xyz_read(const char *path, long nvalues, double *z, int nthreads)
{
    PFX_ARG(path && z && nvalues >= 1, "bad arguments");
    Mapped f;
    if (!f.open_file(path)) {
        set_error("cannot open %s", path);
        return PFX_ERR_ARG;
    }
    const char *b = f.p + header_end(f), *e = f.p + f.n;
    int nt = nthreads > 0 ? nthreads : (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if ((size_t)(e - b) < (1u << 20)) nt = 1;
    // cut at line boundaries; each thread parses its piece into a private vector, pieces are joined in order
    std::vector<const char *> cut(nt + 1, e);
    cut[0] = b;
    for (int t = 1; t < nt; t++) {
        const char *p = b + (size_t)(e - b) * t / nt;
        const char *nl = static_cast<const char *>(memchr(p, '\n', (size_t)(e - p)));
        cut[t] = nl ? nl + 1 : e;
    }
    std::vector<std::vector<double>> part(nt);
    std::vector<long> got(nt, 0);
    std::vector<std::thread> th;
    auto work = [&](int t) {
        const size_t bytes = (size_t)(cut[t + 1] - cut[t]);
        part[t].resize(bytes / 6 + 8);  // a line holds at least "0 0 0\n"
        got[t] = parse_chunk(cut[t], cut[t + 1], part[t].data(), (long)part[t].size());
    };
    for (int t = 1; t < nt; t++) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    long total = 0;
    for (int t = 0; t < nt; t++) {
        if (got[t] < 0) {
            set_error("%s: malformed line (expected three numeric fields)", path);
            return PFX_ERR_ARG;
        }
        total += got[t];
    }
    if (total != nvalues) {
        set_error("%s holds %ld values, expected %ld (nx*ny of the header)", path, total, nvalues);
        return PFX_ERR_ARG;
    }
    long o = 0;
    for (int t = 0; t < nt; t++) {
        memcpy(z + o, part[t].data(), (size_t)got[t] * sizeof(double));
        o += got[t];
    }
    return PFX_OK;
}

int pfx_nan_fill_nearest_host(int device, int nx, int ny, double *grid, long *nfilled)
{
    PFX_ARG(grid && nx >= 1 && ny >= 1, "bad arguments");
    const size_t n = (size_t)nx * ny;
    std::vector<int> todo;
    for (size_t c = 0; c < n; c++)
        if (!std::isfinite(grid[c])) todo.push_back((int)c);
    if (nfilled) *nfilled = (long)todo.size();
    if (todo.empty()) return PFX_OK;
    PFX_ARG(todo.size() < n, "the grid holds no finite value");
    DevGuard guard(device);
    double *d_src = nullptr, *d_dst = nullptr;
    int *d_near = nullptr, *d_todo = nullptr;
    int rc = PFX_OK;
    auto fail = [&](cudaError_t e) {
        set_error("nan fill: %s", cudaGetErrorString(e));
        rc = e == cudaErrorMemoryAllocation ? PFX_ERR_OOM : PFX_ERR_CUDA;
    };
    cudaError_t e = cudaMalloc(&d_src, n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_dst, n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_near, n * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&d_todo, todo.size() * sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpy(d_src, grid, n * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_dst, d_src, n * sizeof(double), cudaMemcpyDeviceToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_todo, todo.data(), todo.size() * sizeof(int), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        column_scan_kernel<<<(ny + 127) / 128, 128>>>(nx, ny, d_src, d_near);
        const long threads = (long)todo.size() * 32;
        fill_kernel<<<(unsigned)((threads + 255) / 256), 256>>>(nx, ny, d_src, d_near, d_todo, (int)todo.size(), d_dst);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(grid, d_dst, n * sizeof(double), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) fail(e);
    cudaFree(d_src);
    cudaFree(d_dst);
    cudaFree(d_near);
    cudaFree(d_todo);
    return rc;
}

int pfx_gaussian_filter_host(int device, int nx, int ny, double sigma, const double *in, double *out)
{
    PFX_ARG(in && out && nx >= 1 && ny >= 1, "bad arguments");
    PFX_ARG(sigma > 0 && std::isfinite(sigma), "sigma must be positive");
    // scipy.ndimage._gaussian_kernel1d: radius = int(truncate * sigma + 0.5), truncate = 4; weights normalised
    const int radius = (int)(4.0 * sigma + 0.5);
    std::vector<double> w(2 * radius + 1);
    double sum = 0.0;
    for (int t = -radius; t <= radius; t++) sum += (w[t + radius] = std::exp(-0.5 / (sigma * sigma) * (double)t * t));
    for (double &v : w) v /= sum;
    DevGuard guard(device);
    const size_t n = (size_t)nx * ny;
    double *d_a = nullptr, *d_b = nullptr, *d_w = nullptr;
    int rc = PFX_OK;
    cudaError_t e = cudaMalloc(&d_a, n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_b, n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&d_w, w.size() * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(d_a, in, n * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_w, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        dim3 grid((ny + 127) / 128, nx);
        gauss_axis_kernel<<<grid, 128>>>(nx, ny, 0, radius, d_w, d_a, d_b);  // scipy filters axis 0 first
        gauss_axis_kernel<<<grid, 128>>>(nx, ny, 1, radius, d_w, d_b, d_a);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, d_a, n * sizeof(double), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) {
        set_error("gaussian filter: %s", cudaGetErrorString(e));
        rc = e == cudaErrorMemoryAllocation ? PFX_ERR_OOM : PFX_ERR_CUDA;
    }
    cudaFree(d_a);
    cudaFree(d_b);
    cudaFree(d_w);
    return rc;
}

}  // extern "C"
