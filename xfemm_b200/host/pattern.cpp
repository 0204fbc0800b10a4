// Sparsity pattern and renumbering (host side, integer work, bit-exact with the reference).
//
//  xfemm_pattern_build      — structure that CBigLinProb grows implicitly through Put()
//                             (cfemm/libfemm/spars.cpp:105-142) while Static2D assembles
//                             (static2d.cpp:345-347, 807-812) and applies Periodicity
//                             (spars.cpp:421-474), produced here in one pass as full CSR.
//  xfemm_renumber_cuthill   — cfemm/libfemm/cuthill.cpp:94-350 (Cuthill–McKee, not reversed)
//  xfemm_renumber_sort_elements — cfemm/libfemm/cuthill.cpp:40-85 (the incomplete comb sort)
#include "../../include/xfemm_b200.h"
#include "../csrc/layout.h"
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <unordered_map>
#include <vector>

namespace xb {
void set_error(const char *fmt, ...);
int element_graph_device(int n, int n_elems, const int *tri, int tri_stride, std::vector<int> &adj, std::vector<int> &rp); // csrc/pattern_device.cu
}

extern "C" {

// timing helper of the pattern functions (XFEMM_TIMING)
struct PatternTicks
{
    bool on = getenv("XFEMM_TIMING") != nullptr;
    double last = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    void operator()(const char *what)
    {
        if (!on) return;
        const double t = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
        fprintf(stderr, "[xfemm timing]   pattern: %s %.3f s\n", what, t - last);
        last = t;
    }
};

// pass 1 on the host: mesh graph (every node coupled to itself and to the nodes it shares an element with) as CSR
// with sorted, unique columns
static int element_graph_host(int n, int n_elems, const int *tri, int tri_stride, std::vector<int> &adj, std::vector<int> &rp,
                              PatternTicks &tick)
{
    // ---- pass 1: element graph as CSR with duplicates, then sort+unique per row.  Integer work over millions of
    // elements: the counting and the scatter run on host threads with relaxed atomics (the order inside a row does
    // not matter, every row is sorted next), the compaction copies rows to their final offsets in parallel.
    std::vector<int> cnt(n + 1, 0);
    {
        std::atomic<int> bad(-1);
        std::vector<std::atomic<int>> acnt(n);
        xb::parallel_for(n, [&](int i0, int i1, int) { for (int i = i0; i < i1; i++) acnt[i].store(1, std::memory_order_relaxed); }); // diagonal
        xb::parallel_for(n_elems, [&](int e0, int e1, int) {
            for (int e = e0; e < e1; e++)
            {
                const int *t = tri + (size_t)e * tri_stride;
                for (int j = 0; j < 3; j++)
                {
                    if (t[j] < 0 || t[j] >= n) { bad.store(e, std::memory_order_relaxed); return; }
                    acnt[t[j]].fetch_add(2, std::memory_order_relaxed);
                }
            }
        });
        if (bad.load() >= 0) { xb::set_error("element %d: node index out of range", bad.load()); return XFEMM_ERR_ARG; }
        for (int i = 0; i < n; i++) cnt[i + 1] = cnt[i] + acnt[i].load(std::memory_order_relaxed);
    }
    tick("count");
    adj.assign((size_t)cnt[n], 0);
    {
        std::vector<std::atomic<int>> pos(n);
        xb::parallel_for(n, [&](int i0, int i1, int) {
            for (int i = i0; i < i1; i++) { adj[cnt[i]] = i; pos[i].store(cnt[i] + 1, std::memory_order_relaxed); }
        });
        xb::parallel_for(n_elems, [&](int e0, int e1, int) {
            for (int e = e0; e < e1; e++)
            {
                const int *t = tri + (size_t)e * tri_stride;
                const int a = t[0], b = t[1], c = t[2];
                const int pa = pos[a].fetch_add(2, std::memory_order_relaxed), pb = pos[b].fetch_add(2, std::memory_order_relaxed),
                          pc = pos[c].fetch_add(2, std::memory_order_relaxed);
                adj[pa] = b; adj[pa + 1] = c;
                adj[pb] = a; adj[pb + 1] = c;
                adj[pc] = a; adj[pc + 1] = b;
            }
        });
    }
    tick("scatter");
    rp.assign(n + 1, 0);
    {
        // sort+unique of every row in parallel, offsets by one serial scan, parallel copy to the compact array
        std::vector<int> len(n);
        xb::parallel_for(n, [&](int i0, int i1, int) {
            for (int i = i0; i < i1; i++)
            {
                int *b = adj.data() + cnt[i], *e = adj.data() + cnt[i + 1];
                std::sort(b, e);
                len[i] = (int)(std::unique(b, e) - b);
            }
        });
        for (int i = 0; i < n; i++) rp[i + 1] = rp[i] + len[i];
        std::vector<int> compact((size_t)rp[n]);
        xb::parallel_for(n, [&](int i0, int i1, int) {
            for (int i = i0; i < i1; i++) memcpy(compact.data() + rp[i], adj.data() + cnt[i], (size_t)len[i] * sizeof(int));
        });
        adj.swap(compact);
    }
    tick("sort + compact");
    return XFEMM_OK;
}

// pass 2 (host): air-gap blocks, explicit couplings and the periodic fill merged into the mesh graph
static int pattern_finish(int n, std::vector<int> &adj, std::vector<int> &rp, int n_age_cells, const int *age_nodes, int n_extra,
                          const int *extra_i, const int *extra_j, int n_pbc, const int *pbc_i, const int *pbc_j, int **rowptr_out,
                          int **colidx_out, unsigned char **fill_out, PatternTicks &tick)
{
    // ---- extras: AGE blocks and explicit couplings (created by assembly), then the periodic
    // fill (symbolic replay of Periodicity in list order; kept apart so that callers can tell
    // which slots exist only if the value-dependent reference code would have created them)
    std::unordered_map<int, std::vector<int>> extra, fill; // row -> additional columns
    auto in_map = [](const std::unordered_map<int, std::vector<int>> &m, int r, int c) -> bool {
        auto f = m.find(r);
        return f != m.end() && std::find(f->second.begin(), f->second.end(), c) != f->second.end();
    };
    auto has = [&](int r, int c) -> bool {
        const int *b = adj.data() + rp[r], *e = adj.data() + rp[r + 1];
        return std::binary_search(b, e, c) || in_map(extra, r, c) || in_map(fill, r, c);
    };
    auto add = [&](std::unordered_map<int, std::vector<int>> &m, int r, int c) {
        if (!has(r, c)) m[r].push_back(c);
        if (r != c && !has(c, r)) m[c].push_back(r);
    };
    for (int k = 0; k < n_age_cells; k++)
    {
        const int *nn = age_nodes + (size_t)k * 10;
        for (int a = 0; a < 10; a++)
        {
            if (nn[a] < 0 || nn[a] >= n) { xb::set_error("AGE cell %d: node out of range", k); return XFEMM_ERR_ARG; }
            for (int b = a; b < 10; b++) add(extra, nn[a], nn[b]);
        }
    }
    for (int k = 0; k < n_extra; k++)
    {
        if (extra_i[k] < 0 || extra_i[k] >= n || extra_j[k] < 0 || extra_j[k] >= n) { xb::set_error("extra coupling %d out of range", k); return XFEMM_ERR_ARG; }
        add(extra, extra_i[k], extra_j[k]);
    }
    std::vector<int> nb;
    for (int k = 0; k < n_pbc; k++)
    {
        const int i = pbc_i[k], j = pbc_j[k];
        if (i < 0 || i >= n || j < 0 || j >= n) { xb::set_error("periodic pair %d out of range", k); return XFEMM_ERR_ARG; }
        nb.clear();
        for (int r : {i, j})
        {
            nb.insert(nb.end(), adj.begin() + rp[r], adj.begin() + rp[r + 1]);
            for (auto *m : {&extra, &fill})
            {
                auto f = m->find(r);
                if (f != m->end()) nb.insert(nb.end(), f->second.begin(), f->second.end());
            }
        }
        std::sort(nb.begin(), nb.end());
        nb.erase(std::unique(nb.begin(), nb.end()), nb.end());
        for (int kk : nb)
        {
            if (kk == i || kk == j) continue;
            add(fill, kk, i);
            add(fill, kk, j);
        }
    }
    tick("extras + periodic fill");
    // ---- merge
    size_t total = adj.size();
    for (auto &kv : extra) total += kv.second.size();
    for (auto &kv : fill) total += kv.second.size();
    int *rowptr = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    int *col = (int *)malloc(sizeof(int) * std::max<size_t>(total, 1));
    unsigned char *isfill = fill_out ? (unsigned char *)calloc(std::max<size_t>(total, 1), 1) : nullptr;
    if (!rowptr || !col || (fill_out && !isfill)) { free(rowptr); free(col); free(isfill); xb::set_error("out of memory"); return XFEMM_ERR_ARG; }
    // row lengths (mesh graph + extras + fill) -> offsets; untouched rows are copied in parallel, the few touched ones
    // (air-gap rings, circuit rows, periodic nodes) are merged serially
    std::vector<char> touched(n, 0);
    for (auto &kv : extra) touched[kv.first] = 1;
    for (auto &kv : fill) touched[kv.first] = 1;
    rowptr[0] = 0;
    for (int i = 0; i < n; i++)
    {
        size_t len = (size_t)(rp[i + 1] - rp[i]);
        if (touched[i])
        {
            auto fe = extra.find(i);
            auto ff = fill.find(i);
            if (fe != extra.end()) len += fe->second.size();
            if (ff != fill.end()) len += ff->second.size();
        }
        rowptr[i + 1] = rowptr[i] + (int)len;
    }
    xb::parallel_for(n, [&](int i0, int i1, int) {
        for (int i = i0; i < i1; i++)
            if (!touched[i]) memcpy(col + rowptr[i], adj.data() + rp[i], (size_t)(rp[i + 1] - rp[i]) * sizeof(int));
    });
    std::vector<int> merged;
    for (int i = 0; i < n; i++)
    {
        if (!touched[i]) continue;
        const int *b = adj.data() + rp[i], *e = adj.data() + rp[i + 1];
        auto fe = extra.find(i);
        auto ff = fill.find(i);
        merged.assign(b, e);
        if (fe != extra.end()) merged.insert(merged.end(), fe->second.begin(), fe->second.end());
        size_t nonfill = merged.size();
        if (ff != fill.end()) merged.insert(merged.end(), ff->second.begin(), ff->second.end());
        // (column, is-fill) pairs; the `has` test above guarantees the three sources are disjoint
        std::vector<std::pair<int, int>> pr(merged.size());
        for (size_t q = 0; q < merged.size(); q++) pr[q] = {merged[q], q >= nonfill ? 1 : 0};
        std::sort(pr.begin(), pr.end());
        size_t w = (size_t)rowptr[i];
        for (auto &x : pr)
        {
            col[w] = x.first;
            if (isfill) isfill[w] = (unsigned char)x.second;
            w++;
        }
    }
    tick("merge");
    if (fill_out) *fill_out = isfill;
    *rowptr_out = rowptr;
    *colidx_out = col;
    return XFEMM_OK;
}

int xfemm_pattern_build(int n_rows, int n_elems, const int *tri, int tri_stride,
                        int n_age_cells, const int *age_nodes,
                        int n_extra, const int *extra_i, const int *extra_j,
                        int n_pbc, const int *pbc_i, const int *pbc_j,
                        int **rowptr_out, int **colidx_out, unsigned char **fill_out)
{
    if (n_rows <= 0 || !rowptr_out || !colidx_out || (n_elems > 0 && (!tri || tri_stride < 3)))
    {
        xb::set_error("xfemm_pattern_build: invalid arguments");
        return XFEMM_ERR_ARG;
    }
    PatternTicks tick;
    std::vector<int> adj, rp;
    int rc = element_graph_host(n_rows, n_elems, tri, tri_stride, adj, rp, tick);
    if (rc != XFEMM_OK) return rc;
    return pattern_finish(n_rows, adj, rp, n_age_cells, age_nodes, n_extra, extra_i, extra_j, n_pbc, pbc_i, pbc_j, rowptr_out,
                          colidx_out, fill_out, tick);
}

// The same with pass 1 on the GPU (csrc/pattern_device.cu: 64-bit (row,column) keys, radix sort, unique) — what
// FSolver uses; needs a CUDA device.
int xfemm_pattern_build_device(int n_rows, int n_elems, const int *tri, int tri_stride,
                               int n_age_cells, const int *age_nodes,
                               int n_extra, const int *extra_i, const int *extra_j,
                               int n_pbc, const int *pbc_i, const int *pbc_j,
                               int **rowptr_out, int **colidx_out, unsigned char **fill_out)
{
    if (n_rows <= 0 || !rowptr_out || !colidx_out || (n_elems > 0 && (!tri || tri_stride < 3)))
    {
        xb::set_error("xfemm_pattern_build_device: invalid arguments");
        return XFEMM_ERR_ARG;
    }
    PatternTicks tick;
    std::vector<int> adj, rp;
    int rc = xb::element_graph_device(n_rows, n_elems, tri, tri_stride, adj, rp);
    if (rc != XFEMM_OK) return rc;
    tick("mesh graph on the device");
    return pattern_finish(n_rows, adj, rp, n_age_cells, age_nodes, n_extra, extra_i, extra_j, n_pbc, pbc_i, pbc_j, rowptr_out,
                          colidx_out, fill_out, tick);
}

// Cuthill–McKee exactly as the reference numbers nodes: adjacency from the .edge list in file
// order, each list stably ordered by neighbour degree (the reference's bubble sort with a strict
// '<' is a stable sort), start at the first node of minimum degree with the "stop scanning once
// a degree-2 node has been seen" shortcut, FIFO numbering, restart on disconnected components.
int xfemm_renumber_cuthill(int n_nodes, int n_edges, const int *edges, int *newnum, int *bandwidth_out)
{
    if (n_nodes <= 0 || n_edges < 0 || !newnum || (n_edges > 0 && !edges))
    {
        xb::set_error("xfemm_renumber_cuthill: invalid arguments");
        return XFEMM_ERR_ARG;
    }
    const int n = n_nodes;
    std::vector<int> deg(n, 0);
    for (int e = 0; e < n_edges; e++)
    {
        int a = edges[2 * e], b = edges[2 * e + 1];
        if (a < 0 || a >= n || b < 0 || b >= n) { xb::set_error("edge %d out of range", e); return XFEMM_ERR_ARG; }
        deg[a]++; deg[b]++;
    }
    std::vector<int> ptr(n + 1, 0);
    for (int i = 0; i < n; i++) ptr[i + 1] = ptr[i] + deg[i];
    std::vector<int> con((size_t)ptr[n]);
    {
        std::vector<int> pos(ptr.begin(), ptr.end() - 1);
        for (int e = 0; e < n_edges; e++)
        {
            int a = edges[2 * e], b = edges[2 * e + 1];
            con[pos[a]++] = b;
            con[pos[b]++] = a;
        }
    }
    xb::parallel_for(n, [&](int i0, int i1, int) {
        for (int i = i0; i < i1; i++)
            std::stable_sort(con.begin() + ptr[i], con.begin() + ptr[i + 1], [&](int a, int b) { return deg[a] < deg[b]; });
    });

    // start node (cuthill.cpp:233-245): the scan stops right after the first index >= 1 at which
    // the running minimum equals 2.
    int start = 0;
    {
        int best = deg[0];
        for (int i = 1; i < n; i++)
        {
            if (deg[i] < best) { best = deg[i]; start = i; }
            if (best == 2) break;
        }
    }
    for (int i = 0; i < n; i++) newnum[i] = -1;
    std::vector<int> queue(n, -1);
    int count = 1, head = 0; // head = position in `queue` of the node being expanded
    newnum[start] = 0;
    queue[0] = start;
    int cur = start;
    if (n > 1)
    {
        do
        {
            for (int s = ptr[cur]; s < ptr[cur + 1]; s++)
            {
                int nbn = con[s];
                if (newnum[nbn] < 0) { newnum[nbn] = count; queue[count] = nbn; count++; }
            }
            head = newnum[cur] + 1;
            if (head < n && queue[head] < 0)
            {
                // component exhausted (cuthill.cpp:269-301): first unvisited node, then the
                // unvisited node of smallest degree (strict '<'), with the same j==2 shortcut
                int pick = -1, best = 0;
                for (int i = 0; i < n; i++)
                    if (newnum[i] < 0) { best = deg[i]; pick = i; break; }
                for (int i = 0; i < n; i++)
                {
                    if (newnum[i] < 0 && deg[i] < best) { best = deg[i]; pick = i; }
                    if (best == 2) break;
                }
                if (pick < 0) break;
                newnum[pick] = count;
                queue[count] = pick;
                count++;
                cur = pick;
            }
            else if (head < n)
                cur = queue[head];
        } while (count < n);
    }
    if (bandwidth_out)
    {
        int wide = 0;
        for (int i = 0; i < n; i++)
            for (int s = ptr[i]; s < ptr[i + 1]; s++)
                wide = std::max(wide, std::abs(newnum[i] - newnum[con[s]]));
        *bandwidth_out = wide + 1;
    }
    return XFEMM_OK;
}

// The reference's element "sort" (cuthill.cpp:40-85) is a comb sort whose loop also ends after
// the first pass without a swap, so the result is only partially ordered; the exact sequence of
// swaps is replayed here on (key, original index) pairs.
int xfemm_renumber_sort_elements(int n_elems, const int *tri, int *order_out)
{
    if (n_elems < 0 || (n_elems > 0 && (!tri || !order_out))) { xb::set_error("invalid arguments"); return XFEMM_ERR_ARG; }
    // (key, index) pairs.  One pass with gap g only ever compares positions of the same residue mod g, so the g chains
    // are independent.
    //  * large gaps: the chains are processed in parallel as contiguous blocks of residues, in place — exactly the
    //    swaps of the serial pass;
    //  * small gaps (few, long chains): a pass over one chain is a bubble pass, i.e. the element carried forward is
    //    the running maximum (ties: the later element, the comparison is a strict '>') and position i-1 receives
    //    min-by-that-rule(carry, element i).  The running maximum is a prefix scan, so the chains are cut into blocks
    //    of consecutive positions: block maxima in parallel, a short serial scan over the blocks, then every block
    //    writes its part of the result (out of place) from its incoming carry.  Same result as the serial pass.
    struct KI { int key, idx; };
    std::vector<KI> bufA(n_elems), bufB;
    KI *a = bufA.data();
    for (int k = 0; k < n_elems; k++) a[k] = {tri[3 * (size_t)k] + tri[3 * (size_t)k + 1] + tri[3 * (size_t)k + 2], k};
    int nthreads = (int)std::thread::hardware_concurrency();
    if (const char *lw = getenv("LOCAL_WORLD_SIZE")) nthreads /= std::max(1, atoi(lw));
    if (const char *e = getenv("XFEMM_HOST_THREADS")) nthreads = atoi(e);
    nthreads = std::max(1, std::min(nthreads, 32));
    int gap = n_elems;
    bool swapped;
    do
    {
        if (gap > 1)
        {
            gap = (int)(((long long)gap * 10) / 13);
            if (gap == 10 || gap == 9) gap = 11;
        }
        const int g = gap;
        if (g <= 0) break;
        std::vector<char> sw(64, 0);
        const long long chain_len = ((long long)n_elems + g - 1) / g; // elements of the longest chain
        const int nblocks = (int)std::min<long long>(nthreads, chain_len / 64);
        if (g < 64 * nthreads && nblocks >= 2 && n_elems > g)
        {
            if (bufB.empty()) bufB.resize(n_elems);
            KI *src = a, *dst = (a == bufA.data()) ? bufB.data() : bufA.data();
            auto blk0 = [&](int b) { return std::min<long long>(n_elems, (chain_len * b / nblocks) * g); }; // first position of block b
            // 1. running maximum of every chain inside every block (ties -> the later element)
            std::vector<KI> bmax((size_t)nblocks * g);
            std::vector<char> bhas((size_t)nblocks * g, 0);
            xb::parallel_for(nblocks, [&](int b0, int b1, int) {
                for (int b = b0; b < b1; b++)
                {
                    KI *mx = bmax.data() + (size_t)b * g;
                    char *has = bhas.data() + (size_t)b * g;
                    const long long p0 = blk0(b), p1 = blk0(b + 1);
                    for (long long base = p0; base < p1; base += g)
                    {
                        const int lim = (int)std::min<long long>(g, p1 - base);
                        for (int r = 0; r < lim; r++)
                        {
                            const KI e = src[base + r];
                            if (!has[r]) { mx[r] = e; has[r] = 1; }
                            else if (!(mx[r].key > e.key)) mx[r] = e;
                        }
                    }
                }
            }, 1);
            // 2. carry entering every block (serial over the blocks, g entries each)
            std::vector<KI> cin((size_t)nblocks * g);
            std::vector<char> chas((size_t)nblocks * g, 0);
            for (int r = 0; r < g; r++)
            {
                bool have = false;
                KI c = {0, 0};
                for (int b = 0; b < nblocks; b++)
                {
                    if (have) { cin[(size_t)b * g + r] = c; chas[(size_t)b * g + r] = 1; }
                    if (bhas[(size_t)b * g + r])
                    {
                        const KI m = bmax[(size_t)b * g + r];
                        if (!have) { c = m; have = true; }
                        else if (!(c.key > m.key)) c = m;
                    }
                }
            }
            // 3. every block writes position (i-1) of its chains for each of its elements i, the last block of a chain
            //    also the final carry
            xb::parallel_for(nblocks, [&](int b0, int b1, int tid) {
                bool s = false;
                std::vector<KI> carry(g);
                std::vector<char> have(g);
                for (int b = b0; b < b1; b++)
                {
                    for (int r = 0; r < g; r++) { have[r] = chas[(size_t)b * g + r]; if (have[r]) carry[r] = cin[(size_t)b * g + r]; }
                    const long long p0 = blk0(b), p1 = blk0(b + 1);
                    for (long long base = p0; base < p1; base += g)
                    {
                        const int lim = (int)std::min<long long>(g, p1 - base);
                        for (int r = 0; r < lim; r++)
                        {
                            const KI e = src[base + r];
                            if (!have[r]) { carry[r] = e; have[r] = 1; continue; } // first element of the chain
                            if (carry[r].key > e.key) { dst[base + r - g] = e; s = true; }
                            else { dst[base + r - g] = carry[r]; carry[r] = e; }
                        }
                    }
                    // chains that end in this block: the carry lands on their last position
                    for (int r = 0; r < g; r++)
                    {
                        if (!have[r]) continue;
                        const long long last = r + ((long long)(n_elems - 1 - r) / g) * g; // last position of chain r
                        if (last >= p0 && last < p1) dst[last] = carry[r];
                    }
                }
                if (s) sw[tid] = 1;
            }, 1);
            a = dst;
        }
        else
        {
            xb::parallel_for(g, [&](int r0, int r1, int tid) {
                bool s = false;
                for (long long base = 0; base + g < n_elems; base += g)
                {
                    const long long hi = std::min<long long>(r1, (long long)n_elems - g - base);
                    KI *x = a + base;
                    for (long long r = r0; r < hi; r++)
                        if (x[r].key > x[r + g].key) { std::swap(x[r], x[r + g]); s = true; }
                }
                if (s) sw[tid] = 1;
            }, 16);
        }
        swapped = false;
        for (char c : sw) swapped = swapped || c;
    } while (gap > 1 && swapped);
    for (int k = 0; k < n_elems; k++) order_out[k] = a[k].idx;
    return XFEMM_OK;
}

} // extern "C"
