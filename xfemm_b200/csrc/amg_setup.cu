// Value-independent set-up of one multigrid coarsening ON THE DEVICE, from the aggregates of a level (amg.cu:
// aggregate(), a sequential greedy pass that stays on a host thread):
//
//   coarse pattern   every fine CSR slot (i,j) gets the key agg[i] * nc + agg[j]; a stable radix sort by key leaves the
//                    fine slots of a coarse slot next to each other in ascending order: the run heads are the coarse CSR
//                    (sorted columns), the run lengths the gather lists of the Galerkin product (gptr/gsrc)
//   centroids        coordinates of a coarse "node" = mean of its members, summed in ascending member order
//   transfer maps    fine internal row -> coarse internal row, and the members of every coarse row (stable sort)
//
// All on the stream the caller passes (the finest level's; the problems' own tables are complete when this runs).
// Array for array the result of the host code in amg.cu (coarse_pattern(), the loops of amg_build()), which remains as
// the reference of the XFEMM_HIERARCHY_CHECK=1 self-test.  At 2 M rows: ~0.1 s of host time and 70 MB of uploads
// become a few milliseconds of sorting.
#include "amg.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <vector>

namespace xb {

static __global__ void k_cp_keys(int n, const int *__restrict__ rowptr, const int *__restrict__ col, const int *__restrict__ agg,
                                 unsigned long long nc, unsigned long long *key, int *slot)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int I = agg[i];
    for (int q = rowptr[i]; q < rowptr[i + 1]; q++)
    {
        const int J = agg[col[q]];
        key[q] = (I >= 0 && J >= 0) ? (unsigned long long)I * nc + (unsigned long long)J : nc * nc; // not in the coarse problem: last
        slot[q] = q;
    }
}

// first position whose key is >= target, for target = R * stride, R = 0..count-1 (row pointers out of sorted keys)
template <class K>
static __global__ void k_lower_bounds(int count, const K *__restrict__ keys, int nkeys, K stride, int *out)
{
    const int R = blockIdx.x * blockDim.x + threadIdx.x;
    if (R >= count) return;
    const K target = (K)R * stride;
    int a = 0, b = nkeys;
    while (a < b) { const int m = (a + b) >> 1; if (keys[m] < target) a = m + 1; else b = m; }
    out[R] = a;
}

static __global__ void k_cp_cols(int nnzc, const unsigned long long *__restrict__ ukey, unsigned long long nc, int *ccol)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nnzc) ccol[t] = (int)(ukey[t] % nc);
}

static __global__ void k_member_keys(int n, const int *__restrict__ agg, unsigned none, unsigned *key, int *idx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    key[i] = agg[i] >= 0 ? (unsigned)agg[i] : none;
    idx[i] = i;
}

static __global__ void k_centroids(int nc, const int *__restrict__ mptr, const int *__restrict__ mem, const float *__restrict__ xy, float *cxy)
{
    const int I = blockIdx.x * blockDim.x + threadIdx.x;
    if (I >= nc) return;
    double sx = 0., sy = 0.;
    const int k0 = mptr[I], k1 = mptr[I + 1];
    for (int k = k0; k < k1; k++) { sx += (double)xy[2 * (size_t)mem[k]]; sy += (double)xy[2 * (size_t)mem[k] + 1]; }
    const int cnt = max(1, k1 - k0);
    cxy[2 * (size_t)I] = (float)(sx / cnt);
    cxy[2 * (size_t)I + 1] = (float)(sy / cnt);
}

static __global__ void k_agg_internal(int n_pad_f, const int *__restrict__ i2c_f, const int *__restrict__ agg, const int *__restrict__ c2i_c,
                                      int *agg_int)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_pad_f) return;
    const int i = i2c_f[r];
    agg_int[r] = (i >= 0 && agg[i] >= 0) ? c2i_c[agg[i]] : -1;
}

template <class T>
static cudaError_t fetch(std::vector<T> &h, const T *d, size_t count, cudaStream_t st)
{
    h.resize(count);
    g_d2h_bytes += (long long)(count * sizeof(T));
    return count ? copy_d2h(h.data(), d, count * sizeof(T), st) : cudaSuccess; // (big ones through the pinned ring)
}

static int bits_for(unsigned long long v) { int b = 1; while (b < 64 && (v >> b)) b++; return b; }

// members of the groups key[i] in 0..ngroups-1 (others last), ascending i inside a group: ptr[ngroups+1], mem[]
static int group_members(int n, const int *d_group /* -1 = none */, int ngroups, DevBuf<int> &ptr, DevBuf<int> &mem, cudaStream_t st)
{
    XB_CUDA(ptr.alloc((size_t)ngroups + 1)); // (the caller's arena, if any; everything else here is temporary)
    NoArena scratch;
    DevBuf<unsigned> ka, kb;
    DevBuf<int> ia;
    DevBuf<unsigned char> tmp;
    XB_CUDA(ka.alloc(n)); XB_CUDA(kb.alloc(n)); XB_CUDA(ia.alloc(n)); XB_CUDA(mem.alloc(n));
    ++g_launch, k_member_keys<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, d_group, (unsigned)ngroups, ka.p, ia.p);
    XB_CUDA(cudaGetLastError());
    size_t bytes = 0;
    const int bits = bits_for((unsigned long long)ngroups);
    XB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, ka.p, kb.p, ia.p, mem.p, n, 0, bits, st));
    XB_CUDA(tmp.alloc(bytes + 256));
    XB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, ka.p, kb.p, ia.p, mem.p, n, 0, bits, st)); // stable
    ++g_launch;
    ++g_launch, k_lower_bounds<unsigned><<<(unsigned)((ngroups + 1 + 255) / 256), 256, 0, st>>>(ngroups + 1, kb.p, n, 1u, ptr.p);
    XB_CUDA(cudaGetLastError());
    XB_CUDA(cudaStreamSynchronize(st)); // the temporaries go out of scope
    return XFEMM_OK;
}

// Coarse pattern of `cur` under the aggregates d_agg (caller numbering of cur, -1 = stays out), nc aggregates.
// Out: host CSR of the coarse problem (sorted columns), device gather lists of the Galerkin product.
int amg_coarse_pattern_device(xfemm_linprob *cur, const int *d_agg, int nc, std::vector<int> &crp, std::vector<int> &ccol,
                              DevBuf<int> &gptr, DevBuf<int> &gsrc, cudaStream_t st)
{
    const int n = cur->n;
    const long long nnz = cur->nnz;
    if (nnz > 0x7fffffffll) { set_error("coarse pattern: %lld entries do not fit 32-bit slots", nnz); return XFEMM_ERR_ARG; }
    DevBuf<unsigned long long> ka, kb;
    DevBuf<int> sa, cnt, d_runs;
    DevBuf<unsigned char> tmp;
    XB_CUDA(ka.alloc((size_t)nnz)); XB_CUDA(kb.alloc((size_t)nnz)); XB_CUDA(sa.alloc((size_t)nnz)); XB_CUDA(gsrc.alloc((size_t)nnz));
    XB_CUDA(cnt.alloc((size_t)nnz + 1)); XB_CUDA(d_runs.alloc(1));
    ++g_launch, k_cp_keys<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, cur->rowptr.p, cur->col.p, d_agg, (unsigned long long)nc, ka.p, sa.p);
    XB_CUDA(cudaGetLastError());
    const int bits = bits_for((unsigned long long)nc * (unsigned long long)nc);
    size_t b1 = 0, b2 = 0, b3 = 0;
    XB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b1, ka.p, kb.p, sa.p, gsrc.p, (int)nnz, 0, bits, st));
    XB_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, b2, kb.p, ka.p, cnt.p, d_runs.p, (int)nnz, st));
    XB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, b3, cnt.p, cnt.p, (int)nnz + 1, st));
    const size_t bytes = std::max(b1, std::max(b2, b3));
    XB_CUDA(tmp.alloc(bytes + 256));
    b1 = b2 = b3 = bytes;
    XB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, b1, ka.p, kb.p, sa.p, gsrc.p, (int)nnz, 0, bits, st)); // stable: ascending fine slot in a run
    ++g_launch;
    // run heads -> ka (unique keys), run lengths -> cnt
    XB_CUDA(cub::DeviceRunLengthEncode::Encode(tmp.p, b2, kb.p, ka.p, cnt.p, d_runs.p, (int)nnz, st));
    ++g_launch;
    int runs = 0;
    unsigned long long last_key = 0;
    XB_CUDA(cudaMemcpyAsync(&runs, d_runs.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    XB_CUDA(cudaStreamSynchronize(st));
    if (runs > 0) XB_CUDA(cudaMemcpy(&last_key, ka.p + (runs - 1), sizeof last_key, cudaMemcpyDeviceToHost));
    const int nnzc = (runs > 0 && last_key == (unsigned long long)nc * (unsigned long long)nc) ? runs - 1 : runs;
    // gather-list offsets: exclusive scan of the run lengths (entry nnzc = number of fine slots that take part)
    XB_CUDA(gptr.alloc((size_t)nnzc + 1));
    XB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, b3, cnt.p, gptr.p, nnzc + 1, st));
    ++g_launch;
    DevBuf<int> d_crp, d_ccol;
    XB_CUDA(d_crp.alloc((size_t)nc + 1)); XB_CUDA(d_ccol.alloc((size_t)std::max(1, nnzc)));
    ++g_launch, k_lower_bounds<unsigned long long><<<(unsigned)((nc + 1 + 255) / 256), 256, 0, st>>>(nc + 1, ka.p, nnzc, (unsigned long long)nc, d_crp.p);
    XB_CUDA(cudaGetLastError());
    if (nnzc > 0)
    {
        ++g_launch, k_cp_cols<<<(unsigned)((nnzc + 255) / 256), 256, 0, st>>>(nnzc, ka.p, (unsigned long long)nc, d_ccol.p);
        XB_CUDA(cudaGetLastError());
    }
    XB_CUDA(fetch(crp, d_crp.p, (size_t)nc + 1, st));
    XB_CUDA(fetch(ccol, d_ccol.p, (size_t)nnzc, st));
    XB_CUDA(cudaStreamSynchronize(st));
    // the gather list keeps only the slots that take part (the tail of the sorted array is the "not in the coarse
    // problem" run; gptr never points into it)
    return XFEMM_OK;
}

// Coordinates of the coarse nodes: c->h_xy (and the device copy c->d_xy) from cur->d_xy
int amg_centroids_device(xfemm_linprob *cur, xfemm_linprob *c, const int *d_agg, cudaStream_t st)
{
    const int nc = c->n;
    NoArena scratch; // nothing here belongs into a level's L2-pinned arena
    DevBuf<int> ptr, mem;
    int rc = group_members(cur->n, d_agg, nc, ptr, mem, st);
    if (rc != XFEMM_OK) return rc;
    XB_CUDA(c->d_xy.alloc(2 * (size_t)nc));
    ++g_launch, k_centroids<<<(unsigned)((nc + 255) / 256), 256, 0, st>>>(nc, ptr.p, mem.p, cur->d_xy.p, c->d_xy.p);
    XB_CUDA(cudaGetLastError());
    XB_CUDA(fetch(c->h_xy, c->d_xy.p, 2 * (size_t)nc, st));
    XB_CUDA(cudaStreamSynchronize(st));
    return XFEMM_OK;
}

// Transfer maps between the internal row orders of cur and c (both layouts built): T.agg, T.mptr, T.mem
int amg_transfer_maps_device(xfemm_linprob *cur, xfemm_linprob *c, const int *d_agg, AmgHierarchy::Transfer &T, cudaStream_t st)
{
    const int npf = cur->lay.n_pad, npc = c->lay.n_pad;
    XB_CUDA(T.agg.alloc((size_t)npf));
    ++g_launch, k_agg_internal<<<(unsigned)((npf + 255) / 256), 256, 0, st>>>(npf, cur->i2c.p, d_agg, c->c2i.p, T.agg.p);
    XB_CUDA(cudaGetLastError());
    DevBuf<int> mem; // (sorted rows: all n_pad of them, the members first)
    int rc = group_members(npf, T.agg.p, npc, T.mptr, mem, st);
    if (rc != XFEMM_OK) return rc;
    // T.mem = the member part (at least one entry: the host version never uploads an empty table)
    std::vector<int> last(1);
    XB_CUDA(cudaMemcpy(last.data(), T.mptr.p + npc, sizeof(int), cudaMemcpyDeviceToHost));
    const int nmem = std::max(1, last[0]);
    XB_CUDA(T.mem.alloc((size_t)nmem));
    XB_CUDA(cudaMemcpyAsync(T.mem.p, mem.p, (size_t)std::min(nmem, npf) * sizeof(int), cudaMemcpyDeviceToDevice, st));
    XB_CUDA(cudaStreamSynchronize(st));
    return XFEMM_OK;
}

} // namespace xb
