// Mesh graph of the sparsity pattern on the GPU (pass 1 of xfemm_pattern_build_device, host/pattern.cpp).
//
// The reference grows the structure implicitly through sorted-list inserts while it assembles
// (cfemm/libfemm/spars.cpp:105-142); here every triangle emits its six ordered node pairs as 64-bit keys
// (row << 32 | column), every node its diagonal, and one radix sort + unique gives the full CSR with sorted columns:
// 26 M keys at 2 M nodes, a few milliseconds instead of the 0.1 s the host threads need.  Sorting runs on CUB
// (set-up step, not the solver's hot loop); results are identical to the host pass (sorted, unique columns).
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <vector>

namespace xb {

__global__ void kp_emit(int n, int n_elems, const int *__restrict__ tri, int stride, unsigned long long *keys, int *bad)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) keys[t] = ((unsigned long long)t << 32) | (unsigned long long)t;
    if (t >= n_elems) return;
    const int *e = tri + t * stride;
    const int a = e[0], b = e[1], c = e[2];
    if (a < 0 || a >= n || b < 0 || b >= n || c < 0 || c >= n) { atomicMin(bad, (int)t); return; } // the host stops before sorting
    unsigned long long *k = keys + n + 6 * t;
    auto key = [](int r, int cc) { return ((unsigned long long)(unsigned)r << 32) | (unsigned long long)(unsigned)cc; };
    k[0] = key(a, b); k[1] = key(a, c); k[2] = key(b, a); k[3] = key(b, c); k[4] = key(c, a); k[5] = key(c, b);
}
__global__ void kp_split(long long m, const unsigned long long *__restrict__ keys, int *col)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < m) col[t] = (int)(keys[t] & 0xffffffffull);
}
__global__ void kp_rowptr(int n, long long m, const unsigned long long *__restrict__ keys, int *rowptr)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    // first key whose row is >= i
    const unsigned long long want = (unsigned long long)i << 32;
    long long lo = 0, hi = m;
    while (lo < hi) { const long long mid = (lo + hi) >> 1; if (keys[mid] < want) lo = mid + 1; else hi = mid; }
    rowptr[i] = (int)lo;
}

int element_graph_device(int n, int n_elems, const int *tri, int tri_stride, std::vector<int> &adj, std::vector<int> &rp)
{
    if (xfemm_device_count() <= 0) { set_error("no CUDA device available: xfemm_pattern_build_device has no CPU fallback"); return XFEMM_ERR_CUDA; }
    const long long total = (long long)n + 6LL * n_elems;
    if (total >= (1LL << 31)) { set_error("pattern too large for 32-bit offsets"); return XFEMM_ERR_ARG; }
    cudaStream_t st = nullptr;
    DevBuf<int> d_tri, d_bad, d_col, d_rp, d_count;
    DevBuf<unsigned long long> k_in, k_out;
    DevBuf<unsigned char> tmp;
    const size_t ntri = (size_t)n_elems * tri_stride;
    if (n_elems > 0) XB_CUDA(d_tri.upload(tri, ntri, st));
    g_h2d_bytes += (long long)ntri * 4;
    XB_CUDA(d_bad.alloc(1));
    const int big = 0x7fffffff;
    XB_CUDA(cudaMemcpyAsync(d_bad.p, &big, sizeof(int), cudaMemcpyHostToDevice, st));
    XB_CUDA(k_in.alloc((size_t)total));
    XB_CUDA(k_out.alloc((size_t)total));
    const long long threads = std::max<long long>(n, n_elems);
    ++g_launch, kp_emit<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(n, n_elems, d_tri.p, tri_stride, k_in.p, d_bad.p);
    int bad = 0;
    XB_CUDA(cudaMemcpyAsync(&bad, d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    XB_CUDA(cudaStreamSynchronize(st));
    if (bad != big) { set_error("element %d: node index out of range", bad); return XFEMM_ERR_ARG; }
    int end_bit = 33;
    while (end_bit < 64 && (1ULL << (end_bit - 32)) < (unsigned long long)n) end_bit++;
    size_t bytes = 0;
    XB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, k_in.p, k_out.p, (int)total, 0, end_bit, st));
    XB_CUDA(tmp.alloc(bytes + 16));
    XB_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, bytes, k_in.p, k_out.p, (int)total, 0, end_bit, st));
    ++g_launch;
    XB_CUDA(d_count.alloc(1));
    size_t bytes2 = 0;
    XB_CUDA(cub::DeviceSelect::Unique(nullptr, bytes2, k_out.p, k_in.p, d_count.p, (int)total, st));
    if (bytes2 > bytes) XB_CUDA(tmp.alloc(bytes2 + 16));
    XB_CUDA(cub::DeviceSelect::Unique(tmp.p, bytes2, k_out.p, k_in.p, d_count.p, (int)total, st));
    ++g_launch;
    int m = 0;
    XB_CUDA(cudaMemcpyAsync(&m, d_count.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    XB_CUDA(cudaStreamSynchronize(st));
    XB_CUDA(d_col.alloc((size_t)m));
    XB_CUDA(d_rp.alloc((size_t)n + 1));
    ++g_launch, kp_split<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(m, k_in.p, d_col.p);
    ++g_launch, kp_rowptr<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(n, m, k_in.p, d_rp.p);
    adj.resize((size_t)m);
    rp.resize((size_t)n + 1);
    XB_CUDA(copy_d2h(adj.data(), d_col.p, (size_t)m * sizeof(int), st));
    XB_CUDA(copy_d2h(rp.data(), d_rp.p, ((size_t)n + 1) * sizeof(int), st));
    XB_CUDA(cudaStreamSynchronize(st));
    XB_CUDA(cudaGetLastError());
    g_d2h_bytes += ((long long)m + n + 1) * 4;
    return XFEMM_OK;
}

} // namespace xb
