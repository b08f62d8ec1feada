// lmc_coords.cu --  CoordinateTransformer.transform_points (CS:214-233) for ONE 4x4
// homogeneous matrix over a flat point buffer, in both summation orders the reference produces:
//
//   LMC_HOMOG_BATCH   (T @ homog.T).T over n >= 2 points runs through dgemm:
//                     fma(T3,1, fma(T2,z, fma(T1,y, T0*x)))   (== the Mode A order with t = T[:,3])
//   LMC_HOMOG_SINGLE  the same expression on ONE point -- what _transform_coordinates does for every
//                     point of every frame (CS:2117-2138) -- runs through a 4-term gemv whose SIMD
//                     kernel multiplies the four lanes and adds them pairwise, products unfused:
//                     (T0*x + T2*z) + (T1*y + T3*1)            (measured against NumPy/OpenBLAS here)
//
// Elementwise and HBM-bound (32 + 32 B/pt f64, 16 + 16 B/pt f32): 256-bit loads / stores, two points
// per thread, grid-stride.  The 4th column (intensity) passes through.
#include "lmc_device.cuh"

namespace lmc {

struct HomogParams { double T[12]; int64_t n; const void* in; void* out; int32_t order; };

__device__ __forceinline__ void homog_apply(const double (&T)[12], int order, Pt& p) {
    const double x = p.x, y = p.y, z = p.z;
    if (order == LMC_HOMOG_BATCH) {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double v = __fma_rn(T[4 * r + 3], 1.0, __fma_rn(T[4 * r + 2], z, __fma_rn(T[4 * r + 1], y, __dmul_rn(T[4 * r], x))));
            (r == 0 ? p.x : r == 1 ? p.y : p.z) = v;
        }
    } else {
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const double v = __dadd_rn(__dadd_rn(__dmul_rn(T[4 * r], x), __dmul_rn(T[4 * r + 2], z)),
                                       __dadd_rn(__dmul_rn(T[4 * r + 1], y), T[4 * r + 3]));
            (r == 0 ? p.x : r == 1 ? p.y : p.z) = v;
        }
    }
}

template <bool F64>
__global__ void __launch_bounds__(256) k_homog(const __grid_constant__ HomogParams P) {
    const int64_t pairs = (P.n + 1) / 2;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < pairs; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = 2 * q;
        const bool vb = i + 1 < P.n;
        Pt a, b{};
        if (vb) load_pair<F64, true>(P.in, i, true, true, a, b); else load_pair<F64, false>(P.in, i, true, false, a, b);
        homog_apply(P.T, P.order, a);
        if (vb) homog_apply(P.T, P.order, b);
        if (vb) store_pair<F64, true>(P.out, i, true, true, a, b); else store_pair<F64, false>(P.out, i, true, false, a, b);
    }
}

cudaError_t launch_homog(bool f64, const void* in, const double* T_host, int32_t order, void* out, int64_t n, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    HomogParams P;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) P.T[4 * r + c] = T_host[4 * r + c];
    P.n = n; P.in = in; P.out = out; P.order = order;
    const int64_t pairs = (n + 1) / 2;
    int64_t blocks = (pairs + 255) / 256;
    const int64_t cap = (int64_t)sms * 8;
    if (blocks > cap) blocks = cap;
    if (f64) k_homog<true><<<(unsigned)blocks, 256, 0, st>>>(P);
    else     k_homog<false><<<(unsigned)blocks, 256, 0, st>>>(P);
    return cudaGetLastError();
}

}  // namespace lmc
