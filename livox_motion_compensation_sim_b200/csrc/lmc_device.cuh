// lmc_device.cuh -- device-side building blocks shared by every kernel of liblmc_b200.so
//
// All floating-point arithmetic that must reproduce the reference bit-for-bit is written with
// explicit __dmul_rn/__dadd_rn/__fma_rn so nvcc can neither contract nor re-associate it; the
// operation orders are the ones NumPy/OpenBLAS execute for the reference's expressions and are
// the same as oracle/lmc_oracle.c (see DESIGN.md "Bit-exact op order").
//
//   LMC = lidar_motion_compensation.py        CS = livox_mid70_complete_simulator.py
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/lmc_b200.h"

namespace lmc {

constexpr int kThreads        = 256;                       // 8 warps per CTA
constexpr int kPairsPerThread = 2;                         // each thread owns 2 pairs of consecutive points
constexpr int kTilePairs      = kThreads * kPairsPerThread;
constexpr int kTile           = 2 * kTilePairs;            // 1024 points per tile
constexpr int kMaxBnd         = 62;                        // frame boundaries cached per tile
constexpr int kSegStride      = 20;                        // doubles per Mode C sample row

enum Mode : int { kRigid = 0, kGyro = 1, kSlerp = 2, kQuantOnly = 3 };

// One launch = one Params block in constant/param space.
struct Params {
    const void*     pts;          // (N,4) f64 or (N) float4
    void*           out;          // same layout, may be null
    const void*     ts;           // int64[N] (f64 layout) | uint32[N] ns offsets (f32 layout)
    const int64_t*  frame_off;    // CSR, n_frames + 1
    const int64_t*  frame_start;  // int64[n_frames] ns
    const double*   pose_Rt;      // Mode A: (n_frames, 12)
    const int64_t*  samp_ts;      // Mode B: imu_ts | Mode C: sample_ts
    const double*   samp_tab;     // Mode B: imu_gyro (S,3) | Mode C: seg (S,20)
    const int32_t*  hold_idx;     // Mode C hold-next (optional)
    int64_t         n_samp;
    int64_t         n_points, p_begin, p_end;
    int32_t         n_frames;
    // fused export epilogues
    uint8_t*        lvx14;
    const uint8_t*  tag;
    int32_t*        las_x;
    int32_t*        las_y;
    int32_t*        las_z;
    uint16_t*       las_int;
    uint32_t*       status;
    double          las_scale[3], las_off[3];
    int32_t         lvx_mode, las_int_mode;
};

struct Pt { double x, y, z, w; };

// ------------------------------------------------------------------------------------------
// 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256).  Points are streamed exactly once, so
// loads bypass L1 allocation; a thread's pair of consecutive points is one (f32) or two (f64)
// fully-used 32-byte sectors.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldg256(const double* p, double& a, double& b, double& c, double& d) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void stg256(double* p, double a, double b, double c, double d) {
    asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
__device__ __forceinline__ void ldg256(const float* p, float (&v)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float (&v)[8]) {
    asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

// ------------------------------------------------------------------------------------------
// Reference operation orders
// ------------------------------------------------------------------------------------------
// sum_k a_k*b_k as OpenBLAS dgemm runs it for the reference's (3x3)@(3xN), N >= 2  (LMC:775)
__device__ __forceinline__ double dot_gemm(double a0, double b0, double a1, double b1, double a2, double b2) {
    return __fma_rn(a2, b2, __fma_rn(a1, b1, __dmul_rn(a0, b0)));
}
// the same through gemv: NumPy takes this route for a 3x1 right-hand side (single-point frame at
// LMC:775) and for every (3,3)@(3,) product (CS:1465)
__device__ __forceinline__ double dot_gemv(double a0, double b0, double a1, double b1, double a2, double b2) {
    return __fma_rn(a2, b2, __fma_rn(a0, b0, __dmul_rn(a1, b1)));
}

// (a2) R p + t with R,t = 12 doubles
__device__ __forceinline__ void rigid_apply(const double (&M)[12], bool single, const Pt& p, Pt& o) {
    if (single) {
        o.x = __dadd_rn(dot_gemv(M[0], p.x, M[1], p.y, M[2], p.z), M[9]);
        o.y = __dadd_rn(dot_gemv(M[3], p.x, M[4], p.y, M[5], p.z), M[10]);
        o.z = __dadd_rn(dot_gemv(M[6], p.x, M[7], p.y, M[8], p.z), M[11]);
    } else {
        o.x = __dadd_rn(dot_gemm(M[0], p.x, M[1], p.y, M[2], p.z), M[9]);
        o.y = __dadd_rn(dot_gemm(M[3], p.x, M[4], p.y, M[5], p.z), M[10]);
        o.z = __dadd_rn(dot_gemm(M[6], p.x, M[7], p.y, M[8], p.z), M[11]);
    }
    o.w = p.w;
}

// (a8) CS:1518-1536  Rx(-rx) @ Ry(-ry) @ Rz(-rz), both 3x3 products in dgemm order, then
// (CS:1465) M @ p in gemv order.  The structural zeros/ones of the factors are kept as literal
// operands so signed zeros and non-finite inputs behave exactly like the reference's dgemm.
__device__ __forceinline__ void gyro_rotate(double ax, double ay, double az, const Pt& p, Pt& o) {
    double sa, ca, sb, cb, sc, cc;
    sincos(-ax, &sa, &ca);
    sincos(-ay, &sb, &cb);
    sincos(-az, &sc, &cc);
    const double Rx[9] = { 1.0, 0.0, 0.0,   0.0, ca, -sa,   0.0, sa, ca };
    const double Ry[9] = { cb, 0.0, sb,   0.0, 1.0, 0.0,   -sb, 0.0, cb };
    const double Rz[9] = { cc, -sc, 0.0,   sc, cc, 0.0,   0.0, 0.0, 1.0 };
    double T[9], M[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            T[3 * r + c] = dot_gemm(Rx[3 * r], Ry[c], Rx[3 * r + 1], Ry[3 + c], Rx[3 * r + 2], Ry[6 + c]);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            M[3 * r + c] = dot_gemm(T[3 * r], Rz[c], T[3 * r + 1], Rz[3 + c], T[3 * r + 2], Rz[6 + c]);
    o.x = dot_gemv(M[0], p.x, M[1], p.y, M[2], p.z);
    o.y = dot_gemv(M[3], p.x, M[4], p.y, M[5], p.z);
    o.z = dot_gemv(M[6], p.x, M[7], p.y, M[8], p.z);
    o.w = p.w;
}

// Mode C per-point evaluation (definition: oracle/lmc_oracle.c::orc_deskew_slerp_f64)
__device__ __forceinline__ void slerp_apply(const double (&s)[kSegStride], double alpha, const Pt& p, Pt& o) {
    const double th = __dmul_rn(alpha, s[15]);
    double sn, cs;
    sincos(th, &sn, &cs);
    const double v = __dsub_rn(1.0, cs);
    const double nx = s[12], ny = s[13], nz = s[14];
    const double c1x = __fma_rn(ny, p.z, -__dmul_rn(nz, p.y));
    const double c1y = __fma_rn(nz, p.x, -__dmul_rn(nx, p.z));
    const double c1z = __fma_rn(nx, p.y, -__dmul_rn(ny, p.x));
    const double c2x = __fma_rn(ny, c1z, -__dmul_rn(nz, c1y));
    const double c2y = __fma_rn(nz, c1x, -__dmul_rn(nx, c1z));
    const double c2z = __fma_rn(nx, c1y, -__dmul_rn(ny, c1x));
    const double x1 = __fma_rn(v, c2x, __fma_rn(sn, c1x, p.x));
    const double y1 = __fma_rn(v, c2y, __fma_rn(sn, c1y, p.y));
    const double z1 = __fma_rn(v, c2z, __fma_rn(sn, c1z, p.z));
    o.x = __dadd_rn(dot_gemm(s[0], x1, s[1], y1, s[2], z1), __fma_rn(alpha, s[16], s[9]));
    o.y = __dadd_rn(dot_gemm(s[3], x1, s[4], y1, s[5], z1), __fma_rn(alpha, s[17], s[10]));
    o.z = __dadd_rn(dot_gemm(s[6], x1, s[7], y1, s[8], z1), __fma_rn(alpha, s[18], s[11]));
    o.w = p.w;
}

// ------------------------------------------------------------------------------------------
// Quantisers
// ------------------------------------------------------------------------------------------
// (a4) LMC:257-259   int(np.clip(v * 1000, -2147483648, 2147483647))
__device__ __forceinline__ int32_t q_mm_clip(double v, uint32_t& fl) {
    double m = __dmul_rn(v, 1000.0);
    if (m != m) { fl |= LMC_FLAG_NAN; return 0; }
    m = fmin(fmax(m, -2147483648.0), 2147483647.0);
    return __double2int_rz(m);
}
// (a4) LMC:266       int(np.clip(i * 255, 0, 255))
__device__ __forceinline__ uint32_t q_refl(double w, uint32_t& fl) {
    double m = __dmul_rn(w, 255.0);
    if (m != m) { fl |= LMC_FLAG_NAN; return 0; }
    m = fmin(fmax(m, 0.0), 255.0);
    return (uint32_t)__double2int_rz(m);
}
// (a9) CS:368-370    int(v * 1000)  -- no clip; '<iii' packing raises outside int32
__device__ __forceinline__ int32_t q_mm_noclip(double v, uint32_t& fl) {
    double m = __dmul_rn(v, 1000.0);
    if (m != m) { fl |= LMC_FLAG_NAN; return 0; }
    if (m >= 2147483648.0 || m <= -2147483649.0) fl |= LMC_FLAG_OVERFLOW;
    return __double2int_rz(m);                      // saturates
}
// (a9) CS:373        struct.pack('<B', point.intensity)
__device__ __forceinline__ uint32_t q_u8_copy(double w, uint32_t& fl) {
    if (w != w) { fl |= LMC_FLAG_NAN; return 0; }
    if (w < 0.0 || w > 255.0) { fl |= LMC_FLAG_OVERFLOW; return w < 0.0 ? 0u : 255u; }
    return (uint32_t)__double2int_rz(w);
}
// (a5)/(a10) laspy: np.round((v - offset) / scale) -> int32   (parity unpinned, see header)
__device__ __forceinline__ int32_t q_las(double v, double scale, double off, uint32_t& fl) {
    double d = __ddiv_rn(__dsub_rn(v, off), scale);
    if (d != d) { fl |= LMC_FLAG_NAN; return 0; }
    if (d >= 2147483647.5 || d < -2147483648.5) fl |= LMC_FLAG_OVERFLOW;
    return __double2int_rn(d);                      // round-half-even, saturates
}
// LMC:961 (w*65535).astype(uint16) | CS:1686 w.astype(uint16): truncate, wrap modulo 2^16
__device__ __forceinline__ uint32_t q_las_intensity(double w, int mode, uint32_t& fl) {
    double m = mode == LMC_LAS_INTENSITY_UNIT ? __dmul_rn(w, 65535.0) : w;
    if (m != m) { fl |= LMC_FLAG_NAN; return 0; }
    if (m >= 9.2e18 || m <= -9.2e18) { fl |= LMC_FLAG_OVERFLOW; return 0; }
    return (uint32_t)(__double2ll_rz(m) & 0xffff);
}

// ------------------------------------------------------------------------------------------
// Warp-cooperative searches over sorted int64 tables (frame_off, imu_ts, sample_ts).
// All 32 lanes must call these together.
// ------------------------------------------------------------------------------------------
// count of elements <= key  (== np.searchsorted(a, key, side='right')); key is warp-uniform.
// 32-ary search: each round probes 32 positions with one coalesced-ish load + ballot.
__device__ __forceinline__ int64_t warp_count_le(const int64_t* __restrict__ a, int64_t n, int64_t key, int lane) {
    int64_t lo = 0, hi = n;                       // answer in [lo, hi]
    while (hi - lo > 32) {
        const int64_t len = hi - lo;
        const int64_t pos = lo + (len * (lane + 1)) / 33;          // 32 probes strictly inside (lo, hi)
        const bool le = __ldg(a + pos) <= key;
        const unsigned m = __ballot_sync(0xffffffffu, le);
        const int c = __popc(m);                                    // probes are sorted: true-prefix
        const int64_t nlo = c > 0 ? lo + (len * c) / 33 + 1 : lo;
        const int64_t nhi = c < 32 ? lo + (len * (c + 1)) / 33 : hi;
        lo = nlo; hi = nhi;
    }
    const int64_t idx = lo + lane;
    const bool le = idx < hi && __ldg(a + idx) <= key;
    return lo + __popc(__ballot_sync(0xffffffffu, le));
}

__device__ __forceinline__ int64_t warp_min_i64(int64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { int64_t t = __shfl_xor_sync(0xffffffffu, v, o); v = t < v ? t : v; }
    return v;
}
__device__ __forceinline__ int64_t warp_max_i64(int64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { int64_t t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
    return v;
}

// per-lane binary search restricted to [lo, hi]: count of a[i] <= key
__device__ __forceinline__ int64_t count_le_in(const int64_t* __restrict__ a, int64_t lo, int64_t hi, int64_t key) {
    while (lo < hi) {
        const int64_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(a + mid) <= key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

}  // namespace lmc
