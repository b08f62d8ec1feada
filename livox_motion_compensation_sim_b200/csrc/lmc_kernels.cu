// lmc_kernels.cu -- "direct" kernels: one 1024-point tile per CTA, 256-bit LDG/STG, fused
// pose lookup + transform + quantise.  The TMA bulk-copy pipelined variant lives in lmc_tma.cu.
//
// Data flow of one CTA (tile = 1024 consecutive points of the merged, frame-major cloud):
//   1. every thread issues its 256-bit point loads (2 pairs of consecutive points) -- no
//      dependence on frame lookup, so the DRAM latency overlaps step 2
//   2. warp 0 finds the frames intersecting the tile: one 32-ary search of the CSR offsets
//      (3 dependent L2 loads for 36 000 frames) + one coalesced read of the following offsets
//   3. per point: frame -> pose (L1 broadcast) -> f64 transform in the reference's op order
//   4. aligned points leave as 256-bit stores; LAS ints as 64-bit SoA stores; LVX 14-byte
//      records are assembled as 7 words per point pair in shared memory (stride 7 words is
//      coprime with 32 banks: conflict-free) and leave as 16-byte coalesced stores
//
//   LMC = lidar_motion_compensation.py        CS = livox_mid70_complete_simulator.py
#include "lmc_device.cuh"

namespace lmc {

struct TileMeta {                 // frames intersecting the tile, in shared memory
    int64_t edge[kMaxBnd + 2];    // frame_off[f_lo .. f_lo + nb + 1]
    int32_t f_lo;                 // frame of the tile's first point
    int32_t nb;                   // boundaries inside the tile (clamped use: nb <= kMaxBnd)
    int32_t overflow;             // more than kMaxBnd boundaries: per-point global search
};

// Warp 0: fill TileMeta for points [first, last].
__device__ __forceinline__ void tile_meta(const Params& P, int64_t first, int64_t last, TileMeta& tm, int lane) {
    const int64_t F = P.n_frames;
    const int64_t cnt = warp_count_le(P.frame_off, F + 1, first, lane);   // >= 1 since frame_off[0] = 0
    const int64_t f_lo = cnt - 1;
    // offsets after f_lo: collect those <= last (boundaries inside the tile) + the closing edge
    int nb = 0; bool done = false; int overflow = 0;
    if (lane == 0) tm.edge[0] = __ldg(P.frame_off + f_lo);
    for (int64_t j0 = f_lo + 1; !done; j0 += 32) {
        const int64_t j = j0 + lane;
        const int64_t v = j <= F ? __ldg(P.frame_off + j) : INT64_MAX;
        const unsigned m = __ballot_sync(0xffffffffu, v <= last);       // true-prefix (sorted)
        const int c = __popc(m);
        // store boundaries and the first edge beyond them
        if (lane <= c && nb + lane < kMaxBnd + 1) tm.edge[1 + nb + lane] = v;
        nb += c;
        done = c < 32;
        if (nb > kMaxBnd) { overflow = 1; done = true; }
    }
    if (lane == 0) { tm.f_lo = (int32_t)f_lo; tm.nb = nb; tm.overflow = overflow; }
}

// frame of point p (global index) + whether that frame holds exactly one point
__device__ __forceinline__ int32_t frame_of(const Params& P, const TileMeta& tm, int64_t p, bool& single) {
    if (!tm.overflow) {
        int lo = 0, hi = tm.nb;                     // count of edge[1..nb] <= p
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (tm.edge[1 + mid] <= p) lo = mid + 1; else hi = mid; }
        single = (tm.edge[lo + 1] - tm.edge[lo]) == 1;
        return tm.f_lo + lo;
    }
    const int64_t f = count_le_in(P.frame_off, 0, (int64_t)P.n_frames + 1, p) - 1;
    single = (__ldg(P.frame_off + f + 1) - __ldg(P.frame_off + f)) == 1;
    return (int32_t)f;
}

template <bool F64> struct Layout;
template <> struct Layout<true>  { using T = double; };
template <> struct Layout<false> { using T = float;  };

// ---- point pair load / store ------------------------------------------------------------
template <bool F64>
__device__ __forceinline__ void load_pair(const void* base, int64_t p, bool va, bool vb, Pt& a, Pt& b) {
    if constexpr (F64) {
        const double* src = reinterpret_cast<const double*>(base) + 4 * p;
        if (va) ldg256(src, a.x, a.y, a.z, a.w);
        if (vb) ldg256(src + 4, b.x, b.y, b.z, b.w);
    } else {
        const float* src = reinterpret_cast<const float*>(base) + 4 * p;
        if (va && vb) {
            float v[8];
            ldg256(src, v);
            a = { (double)v[0], (double)v[1], (double)v[2], (double)v[3] };
            b = { (double)v[4], (double)v[5], (double)v[6], (double)v[7] };
        } else {
            if (va) { const float4 v = __ldg(reinterpret_cast<const float4*>(src));     a = { (double)v.x, (double)v.y, (double)v.z, (double)v.w }; }
            if (vb) { const float4 v = __ldg(reinterpret_cast<const float4*>(src) + 1); b = { (double)v.x, (double)v.y, (double)v.z, (double)v.w }; }
        }
    }
}

template <bool F64>
__device__ __forceinline__ void store_pair(void* base, int64_t p, bool va, bool vb, const Pt& a, const Pt& b) {
    if constexpr (F64) {
        double* dst = reinterpret_cast<double*>(base) + 4 * p;
        if (va) stg256(dst, a.x, a.y, a.z, a.w);
        if (vb) stg256(dst + 4, b.x, b.y, b.z, b.w);
    } else {
        float* dst = reinterpret_cast<float*>(base) + 4 * p;
        if (va && vb) {
            const float v[8] = { (float)a.x, (float)a.y, (float)a.z, (float)a.w, (float)b.x, (float)b.y, (float)b.z, (float)b.w };
            stg256(dst, v);
        } else {
            if (va) *reinterpret_cast<float4*>(dst)     = make_float4((float)a.x, (float)a.y, (float)a.z, (float)a.w);
            if (vb) *reinterpret_cast<float4*>(dst + 4) = make_float4((float)b.x, (float)b.y, (float)b.z, (float)b.w);
        }
    }
}

// ---- fused kernel --------------------------------------------------------------------------
template <bool F64, int MODE>
__global__ void __launch_bounds__(kThreads) k_fused(const __grid_constant__ Params P)
{
    __shared__ __align__(16) uint32_t s_lvx[kTilePairs * 7];      // 1024 x 14 B records
    __shared__ TileMeta s_tm;

    const int tid = threadIdx.x, lane = tid & 31;
    const int64_t tile0 = (P.p_begin / kTile) * kTile;            // tiles are aligned in GLOBAL index space
    const int64_t base  = tile0 + (int64_t)blockIdx.x * kTile;
    const int64_t lim_lo = base > P.p_begin ? base : P.p_begin;
    const int64_t lim_hi = base + kTile < P.p_end ? base + kTile : P.p_end;
    if (lim_lo >= lim_hi) return;

    // ---- 1. loads --------------------------------------------------------------------------
    Pt in[kPairsPerThread][2];
    bool valid[kPairsPerThread][2];
    int64_t tsv[kPairsPerThread][2];
    uint32_t tagv[kPairsPerThread];
#pragma unroll
    for (int j = 0; j < kPairsPerThread; ++j) {
        const int64_t p = base + 2 * (tid + j * kThreads);
        const bool va = p >= lim_lo && p < lim_hi, vb = p + 1 >= lim_lo && p + 1 < lim_hi;
        valid[j][0] = va; valid[j][1] = vb;
        in[j][0] = in[j][1] = Pt{ 0.0, 0.0, 0.0, 0.0 };
        load_pair<F64>(P.pts, p, va, vb, in[j][0], in[j][1]);
        tsv[j][0] = tsv[j][1] = 0;
        if constexpr (MODE == kGyro || MODE == kSlerp) {
            if (P.ts != nullptr) {
                if constexpr (F64) {
                    const int64_t* t = reinterpret_cast<const int64_t*>(P.ts) + p;
                    if (va && vb) { const longlong2 v = __ldg(reinterpret_cast<const longlong2*>(t)); tsv[j][0] = v.x; tsv[j][1] = v.y; }
                    else { if (va) tsv[j][0] = __ldg(t); if (vb) tsv[j][1] = __ldg(t + 1); }
                } else {
                    const uint32_t* t = reinterpret_cast<const uint32_t*>(P.ts) + p;
                    if (va && vb) { const uint2 v = __ldg(reinterpret_cast<const uint2*>(t)); tsv[j][0] = v.x; tsv[j][1] = v.y; }
                    else { if (va) tsv[j][0] = __ldg(t); if (vb) tsv[j][1] = __ldg(t + 1); }
                }
            }
        }
        tagv[j] = 0;
        if (P.lvx14 != nullptr && P.tag != nullptr && P.lvx_mode == LMC_LVX2_OF_OUTPUT) {
            if (va) tagv[j] |= __ldg(P.tag + p);
            if (vb) tagv[j] |= (uint32_t)__ldg(P.tag + p + 1) << 8;
        }
    }

    // ---- 2. frames intersecting the tile ------------------------------------------------------
    if constexpr (MODE != kQuantOnly) {
        if (tid < 32) tile_meta(P, lim_lo, lim_hi - 1, s_tm, lane);
        __syncthreads();
    }

    // ---- 3. compute ---------------------------------------------------------------------------
    Pt outp[kPairsPerThread][2];
    uint32_t fl = 0;

    if constexpr (MODE == kRigid) {
        double M[12];
        int32_t fc = -1;
#pragma unroll
        for (int j = 0; j < kPairsPerThread; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                outp[j][h] = in[j][h];
                if (!valid[j][h]) continue;
                bool single;
                const int32_t f = frame_of(P, s_tm, base + 2 * (tid + j * kThreads) + h, single);
                if (f != fc) {
                    const double2* pr = reinterpret_cast<const double2*>(P.pose_Rt + 12 * (int64_t)f);
#pragma unroll
                    for (int q = 0; q < 6; ++q) { const double2 v = __ldg(pr + q); M[2 * q] = v.x; M[2 * q + 1] = v.y; }
                    fc = f;
                }
                rigid_apply(M, single, in[j][h], outp[j][h]);
            }
    } else if constexpr (MODE == kGyro) {
        // (a6)/(a7): per-point bracket in imu_ts, gyro lerp, small-angle rotation
        const int64_t S = P.n_samp;
        int32_t fr[kPairsPerThread][2];
        int64_t kmin = INT64_MAX, kmax = INT64_MIN;
#pragma unroll
        for (int j = 0; j < kPairsPerThread; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                fr[j][h] = 0;
                if (!valid[j][h]) continue;
                bool single;
                fr[j][h] = frame_of(P, s_tm, base + 2 * (tid + j * kThreads) + h, single);
                if constexpr (!F64) tsv[j][h] += __ldg(P.frame_start + fr[j][h]);
                kmin = tsv[j][h] < kmin ? tsv[j][h] : kmin;
                kmax = tsv[j][h] > kmax ? tsv[j][h] : kmax;
            }
        int64_t wlo = 0, whi = 0;
        if (S > 0) {
            kmin = warp_min_i64(kmin); kmax = warp_max_i64(kmax);
            if (kmin <= kmax) {                                   // warp has at least one valid point
                wlo = warp_count_le(P.samp_ts, S, kmin, lane);
                whi = kmin == kmax ? wlo : warp_count_le(P.samp_ts, S, kmax, lane);
            }
        }
#pragma unroll
        for (int j = 0; j < kPairsPerThread; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                outp[j][h] = in[j][h];
                if (!valid[j][h] || S == 0) continue;            // CS:1439-1440: no IMU data -> unchanged
                const int64_t t = tsv[j][h];
                const int64_t k = count_le_in(P.samp_ts, wlo, whi, t) - 1;   // index of `before`
                double g0, g1, g2;
                if (k < 0 || k >= S - 1) {                       // CS:1495-1496: clamp to the existing end sample
                    const double* g = P.samp_tab + 3 * (k < 0 ? 0 : S - 1);
                    g0 = __ldg(g); g1 = __ldg(g + 1); g2 = __ldg(g + 2);
                } else {
                    const int64_t tb = __ldg(P.samp_ts + k), ta = __ldg(P.samp_ts + k + 1);
                    const double alpha = __ddiv_rn((double)(t - tb), (double)(ta - tb));     // CS:1503
                    const double* gb = P.samp_tab + 3 * k;
                    const double b0 = __ldg(gb), b1 = __ldg(gb + 1), b2 = __ldg(gb + 2);
                    const double a0 = __ldg(gb + 3), a1 = __ldg(gb + 4), a2 = __ldg(gb + 5);
                    g0 = __dadd_rn(b0, __dmul_rn(alpha, __dsub_rn(a0, b0)));                 // CS:1507-1509
                    g1 = __dadd_rn(b1, __dmul_rn(alpha, __dsub_rn(a1, b1)));
                    g2 = __dadd_rn(b2, __dmul_rn(alpha, __dsub_rn(a2, b2)));
                }
                const double dt = __dmul_rn((double)(t - __ldg(P.frame_start + fr[j][h])), 1e-9);   // CS:1454
                gyro_rotate(__dmul_rn(g0, dt), __dmul_rn(g1, dt), __dmul_rn(g2, dt), in[j][h], outp[j][h]);
            }
    } else if constexpr (MODE == kSlerp) {
        const int64_t S = P.n_samp;
        const bool hold = P.hold_idx != nullptr;
        int32_t fr[kPairsPerThread][2];
        int64_t kmin = INT64_MAX, kmax = INT64_MIN;
#pragma unroll
        for (int j = 0; j < kPairsPerThread; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                fr[j][h] = 0;
                if (!valid[j][h]) continue;
                bool single;
                fr[j][h] = frame_of(P, s_tm, base + 2 * (tid + j * kThreads) + h, single);
                if constexpr (!F64) { if (!hold) tsv[j][h] += __ldg(P.frame_start + fr[j][h]); }
                kmin = tsv[j][h] < kmin ? tsv[j][h] : kmin;
                kmax = tsv[j][h] > kmax ? tsv[j][h] : kmax;
            }
        int64_t wlo = 0, whi = 0;
        if (!hold) {
            kmin = warp_min_i64(kmin); kmax = warp_max_i64(kmax);
            if (kmin <= kmax) {
                wlo = warp_count_le(P.samp_ts, S, kmin, lane);
                whi = kmin == kmax ? wlo : warp_count_le(P.samp_ts, S, kmax, lane);
            }
        }
        double s[kSegStride];
        int64_t kc = -1;
#pragma unroll
        for (int j = 0; j < kPairsPerThread; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                outp[j][h] = in[j][h];
                if (!valid[j][h]) continue;
                int64_t k; double alpha = 0.0;
                if (hold) k = __ldg(P.hold_idx + fr[j][h]);
                else {
                    const int64_t t = tsv[j][h];
                    k = count_le_in(P.samp_ts, wlo, whi, t) - 1;
                    if (k < 0) k = 0;
                    else if (k >= S - 1) k = S - 1;
                    else {
                        const int64_t tb = __ldg(P.samp_ts + k), ta = __ldg(P.samp_ts + k + 1);
                        alpha = __ddiv_rn((double)(t - tb), (double)(ta - tb));
                    }
                }
                if (k != kc) {
                    const double2* sr = reinterpret_cast<const double2*>(P.samp_tab + kSegStride * k);
#pragma unroll
                    for (int q = 0; q < kSegStride / 2; ++q) { const double2 v = __ldg(sr + q); s[2 * q] = v.x; s[2 * q + 1] = v.y; }
                    kc = k;
                }
                slerp_apply(s, alpha, in[j][h], outp[j][h]);
            }
    } else {
#pragma unroll
        for (int j = 0; j < kPairsPerThread; ++j) { outp[j][0] = in[j][0]; outp[j][1] = in[j][1]; }
    }

    // ---- 4. stores + export epilogues ---------------------------------------------------------
#pragma unroll
    for (int j = 0; j < kPairsPerThread; ++j) {
        const int q = tid + j * kThreads;
        const int64_t p = base + 2 * q;
        const bool va = valid[j][0], vb = valid[j][1];
        if (P.out != nullptr) store_pair<F64>(P.out, p, va, vb, outp[j][0], outp[j][1]);

        if (P.las_x != nullptr) {
            int32_t X[2], Y[2], Z[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                X[h] = Y[h] = Z[h] = 0;
                if (!valid[j][h]) continue;
                X[h] = q_las(outp[j][h].x, P.las_scale[0], P.las_off[0], fl);
                Y[h] = q_las(outp[j][h].y, P.las_scale[1], P.las_off[1], fl);
                Z[h] = q_las(outp[j][h].z, P.las_scale[2], P.las_off[2], fl);
            }
            if (va && vb) {
                *reinterpret_cast<int2*>(P.las_x + p) = make_int2(X[0], X[1]);
                *reinterpret_cast<int2*>(P.las_y + p) = make_int2(Y[0], Y[1]);
                *reinterpret_cast<int2*>(P.las_z + p) = make_int2(Z[0], Z[1]);
            } else {
                if (va) { P.las_x[p] = X[0]; P.las_y[p] = Y[0]; P.las_z[p] = Z[0]; }
                if (vb) { P.las_x[p + 1] = X[1]; P.las_y[p + 1] = Y[1]; P.las_z[p + 1] = Z[1]; }
            }
        }
        if (P.las_int != nullptr) {
            uint32_t i0 = 0, i1 = 0;
            if (va) i0 = q_las_intensity(outp[j][0].w, P.las_int_mode, fl);
            if (vb) i1 = q_las_intensity(outp[j][1].w, P.las_int_mode, fl);
            if (va && vb) *reinterpret_cast<uint32_t*>(P.las_int + p) = i0 | (i1 << 16);
            else { if (va) P.las_int[p] = (uint16_t)i0; if (vb) P.las_int[p + 1] = (uint16_t)i1; }
        }
        if (P.lvx14 != nullptr) {
            uint32_t x[2], y[2], z[2], rt[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                x[h] = y[h] = z[h] = rt[h] = 0;
                if (!valid[j][h]) continue;
                if (P.lvx_mode == LMC_LVX_TYPE2_OF_INPUT) {           // LMC:252-272 on the raw point
                    const Pt& s = MODE == kQuantOnly ? outp[j][h] : in[j][h];
                    x[h] = (uint32_t)q_mm_clip(s.x, fl); y[h] = (uint32_t)q_mm_clip(s.y, fl); z[h] = (uint32_t)q_mm_clip(s.z, fl);
                    rt[h] = q_refl(s.w, fl);                          // tag byte = 0
                } else {                                              // CS:365-374 on the compensated point
                    const Pt& s = outp[j][h];
                    x[h] = (uint32_t)q_mm_noclip(s.x, fl); y[h] = (uint32_t)q_mm_noclip(s.y, fl); z[h] = (uint32_t)q_mm_noclip(s.z, fl);
                    rt[h] = q_u8_copy(s.w, fl) | (((tagv[j] >> (8 * h)) & 0xffu) << 8);
                }
            }
            uint32_t* w = s_lvx + 7 * q;                               // 28 B per pair, conflict-free stride
            w[0] = x[0]; w[1] = y[0]; w[2] = z[0];
            w[3] = rt[0] | (x[1] << 16);
            w[4] = (x[1] >> 16) | (y[1] << 16);
            w[5] = (y[1] >> 16) | (z[1] << 16);
            w[6] = (z[1] >> 16) | (rt[1] << 16);
        }
    }

    if (P.lvx14 != nullptr) {
        __syncthreads();
        // copy bytes [b0, b1) of the tile's record block: 16-byte body, byte-wise ragged ends
        const int b0 = (int)(lim_lo - base) * 14, b1 = (int)(lim_hi - base) * 14;
        uint8_t* g = P.lvx14 + 14 * base;
        const uint8_t* s = reinterpret_cast<const uint8_t*>(s_lvx);
        int a0 = (b0 + 15) & ~15; if (a0 > b1) a0 = b1;
        int a1 = b1 & ~15;        if (a1 < a0) a1 = a0;
        for (int i = a0 / 16 + tid; i < a1 / 16; i += kThreads)
            reinterpret_cast<uint4*>(g)[i] = reinterpret_cast<const uint4*>(s)[i];
        for (int i = b0 + tid; i < a0; i += kThreads) g[i] = s[i];
        for (int i = a1 + tid; i < b1; i += kThreads) g[i] = s[i];
    }
    if (fl != 0 && P.status != nullptr) atomicOr(P.status, fl);
}

// (a1) LMC:802-812: one thread per frame
__global__ void k_pose_lookup(const double* __restrict__ traj_t, int64_t n_t, const double* __restrict__ traj_Rt,
                              const double* __restrict__ frame_t, int32_t n_frames,
                              double* __restrict__ pose_Rt, int32_t* __restrict__ pose_idx)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frames) return;
    const double t = frame_t[f];
    int64_t lo = 0, hi = n_t;                       // np.searchsorted side='left'
    while (lo < hi) { const int64_t mid = lo + ((hi - lo) >> 1); if (traj_t[mid] < t) lo = mid + 1; else hi = mid; }
    if (lo > n_t - 1) lo = n_t - 1;
    if (lo < 0) lo = 0;
    const double2* src = reinterpret_cast<const double2*>(traj_Rt + 12 * lo);
    double2* dst = reinterpret_cast<double2*>(pose_Rt + 12 * (int64_t)f);
#pragma unroll
    for (int q = 0; q < 6; ++q) dst[q] = __ldg(src + q);
    if (pose_idx) pose_idx[f] = (int32_t)lo;
}

// ---- launchers -------------------------------------------------------------------------------
template <bool F64, int MODE>
static cudaError_t launch_fused(const Params& P, cudaStream_t st) {
    if (P.p_end <= P.p_begin) return cudaSuccess;
    const int64_t tile0 = (P.p_begin / kTile) * kTile;
    const int64_t tiles = (P.p_end - tile0 + kTile - 1) / kTile;
    if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
    k_fused<F64, MODE><<<(unsigned)tiles, kThreads, 0, st>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_direct(bool f64, int mode, const Params& P, cudaStream_t st) {
    switch (mode) {
    case kRigid:     return f64 ? launch_fused<true, kRigid>(P, st)     : launch_fused<false, kRigid>(P, st);
    case kGyro:      return f64 ? launch_fused<true, kGyro>(P, st)      : launch_fused<false, kGyro>(P, st);
    case kSlerp:     return f64 ? launch_fused<true, kSlerp>(P, st)     : launch_fused<false, kSlerp>(P, st);
    case kQuantOnly: return f64 ? launch_fused<true, kQuantOnly>(P, st) : launch_fused<false, kQuantOnly>(P, st);
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_pose_lookup(const double* traj_t, int64_t n_t, const double* traj_Rt, const double* frame_t,
                               int32_t n_frames, double* pose_Rt, int32_t* pose_idx, cudaStream_t st) {
    if (n_frames <= 0) return cudaSuccess;
    k_pose_lookup<<<(n_frames + 127) / 128, 128, 0, st>>>(traj_t, n_t, traj_Rt, frame_t, n_frames, pose_Rt, pose_idx);
    return cudaGetLastError();
}

}  // namespace lmc
