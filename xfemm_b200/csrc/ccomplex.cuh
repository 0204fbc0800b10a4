// double2 complex arithmetic shared by the complex solver (clinprob.cu) and its Newton-AC extension (cnewton.cu).
#pragma once
#include "common.cuh"

namespace xb {

typedef double2 cd;
__device__ __forceinline__ cd cmul(cd a, cd b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ cd cadd(cd a, cd b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cd csub(cd a, cd b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cd cscale(cd a, double s) { return make_double2(a.x * s, a.y * s); }
__device__ __forceinline__ cd cconj(cd a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ cd cfma(cd a, cd b, cd c) // a*b + c
{
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}
__device__ __forceinline__ cd cdiv(cd a, cd b)
{
    // Smith's algorithm
    if (fabs(b.x) >= fabs(b.y))
    {
        double t = b.y / b.x, d = b.x + b.y * t;
        return make_double2((a.x + a.y * t) / d, (a.y - a.x * t) / d);
    }
    double t = b.x / b.y, d = b.x * t + b.y;
    return make_double2((a.x * t + a.y) / d, (a.y * t - a.x) / d);
}
// SetValue (cspars.cpp:482-537): row k is scanned for column i when k lies in [i-bdw, min(i+bdw,NumNodes))
// or k is a circuit row (k >= NumNodes, always scanned)
__device__ __forceinline__ bool cband(int k, int i, int bdw, int n_nodes)
{
    if (k >= n_nodes) return true;
    if (bdw == 0) return true;
    int lst = i + bdw;
    if (lst > n_nodes) lst = n_nodes;
    return k >= i - bdw && k < lst;
}

} // namespace xb
