// lmc_las.cu --  a complete LAS 1.2 / point-format-3 file image built on the device:
// what `save_las` (LMC:950-963) / `_export_las` (CS:1671-1698) obtain from laspy.
//
// PARITY UNPINNED: laspy is neither vendored nor installed, so there are no reference bytes to
// compare with.  The file follows the public LAS 1.2 specification: 227-byte public header block,
// no VLRs, 34-byte PF3 records {X Y Z int32, intensity u16, return/flag byte, classification,
// scan angle, user data, point source id u16, gps_time f64, R G B u16}.  X/Y/Z are the same
// rint((v - offset) / scale) integers as the fused LAS epilogue (bit-exact against the restatement),
// intensity follows the two call sites (LMC:961 unit scale, CS:1686 raw), every field the reference
// never sets stays 0 (as laspy's zero-initialised record array would leave it), and the header's
// min / max are the scaled extremes X*scale + offset (what laspy's update_header computes), reduced on
// the GPU with integer atomics.
//
// Two launches: k_las_records (records + min/max reduction; the 227-byte data offset makes every
// record start at an odd address, so each CTA assembles its contiguous byte range in shared memory at
// the destination's 16-byte phase and hands the aligned body to the TMA engine as one bulk store), then k_las_header.
#include "lmc_device.cuh"

namespace lmc {

constexpr int kLasHeader  = 227;
constexpr int kLasRec     = 34;
constexpr int kLasTile    = 512;                    // points per CTA
constexpr int kLasThreads = 256;
constexpr int kLasImg     = ((kLasTile * kLasRec + 32 + 15) / 16) * 16;

__device__ __forceinline__ void put_le(uint8_t* p, uint64_t v, int nbytes) {
#pragma unroll
    for (int k = 0; k < 8; ++k) if (k < nbytes) p[k] = (uint8_t)(v >> (8 * k));
}

template <bool F64>
__global__ void __launch_bounds__(kLasThreads) k_las_records(const __grid_constant__ LasParams L) {
    __shared__ __align__(16) uint8_t s_img[kLasImg];
    __shared__ int32_t s_mm[6];
    const int tid = threadIdx.x;
    // points [p_begin, n) of the cloud; L.out is the (possibly virtual) address of file byte 0 (see lmc_las_pf3_records_*)
    const int64_t first = L.p_begin + (int64_t)blockIdx.x * kLasTile;
    const int cnt = (int)min((int64_t)kLasTile, L.n - first);
    const int64_t dst0 = kLasHeader + first * kLasRec;
    const int phase = (int)(reinterpret_cast<uintptr_t>(L.out + dst0) & 15);
    uint8_t* img = s_img + phase;
    if (tid < 6) s_mm[tid] = (tid & 1) ? INT32_MIN : INT32_MAX;
    __syncthreads();
    uint32_t fl = 0;
    int32_t mn[3] = { INT32_MAX, INT32_MAX, INT32_MAX }, mx[3] = { INT32_MIN, INT32_MIN, INT32_MIN };
    for (int j = tid; j < cnt; j += kLasThreads) {
        Pt p;
        const int64_t i = first + j;
        if constexpr (F64) { const double* s = reinterpret_cast<const double*>(L.pts) + 4 * i; ldg256(s, p.x, p.y, p.z, p.w); }
        else { const float4 v = __ldg(reinterpret_cast<const float4*>(L.pts) + i); p = Pt{ (double)v.x, (double)v.y, (double)v.z, (double)v.w }; }
        // (the always-exact route: this kernel is bound by its 34-byte records, not by FP64 -- the tie test of q_las costs it 2 %)
        const int2 qx = q_las_exact_body(p.x, L.scale[0], L.rcp[0], L.off[0]), qy = q_las_exact_body(p.y, L.scale[1], L.rcp[1], L.off[1]),
                   qz = q_las_exact_body(p.z, L.scale[2], L.rcp[2], L.off[2]);
        const int32_t X = qx.x, Y = qy.x, Z = qz.x;
        fl |= (uint32_t)(qx.y | qy.y | qz.y);
        const uint32_t I = q_las_intensity(p.w, L.intensity_mode, fl);
        mn[0] = min(mn[0], X); mx[0] = max(mx[0], X); mn[1] = min(mn[1], Y); mx[1] = max(mx[1], Y); mn[2] = min(mn[2], Z); mx[2] = max(mx[2], Z);
        // the record starts at an odd address (227 + 34 j): one byte, sixteen aligned halfwords, one byte
        uint8_t* r = img + j * kLasRec;
        const uint64_t gb = (uint64_t)__double_as_longlong(L.gps_time ? __ldg(L.gps_time + i) : 0.0);
        // bytes 0-11 X Y Z | 12-13 intensity | 14-19 return byte, classification, scan angle, user data, point source id = 0
        // | 20-27 gps_time | 28-33 R G B = 0
        const uint32_t W[9] = { (uint32_t)X, (uint32_t)Y, (uint32_t)Z, I & 0xffffu, 0u, (uint32_t)gb, (uint32_t)(gb >> 32), 0u, 0u };
        r[0] = (uint8_t)W[0];
        uint16_t* h = reinterpret_cast<uint16_t*>(r + 1);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const int b = 1 + 2 * k;                                  // first byte of this halfword
            h[k] = (b & 3) == 1 ? (uint16_t)(W[b >> 2] >> 8) : (uint16_t)((W[b >> 2] >> 24) | (W[(b >> 2) + 1] << 8));
        }
        r[33] = 0;
    }
    // block min / max -> one atomic per value per CTA
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { mn[c] = min(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o)); mx[c] = max(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o)); }
        if ((tid & 31) == 0) { atomicMin(&s_mm[2 * c], mn[c]); atomicMax(&s_mm[2 * c + 1], mx[c]); }
    }
    __syncthreads();
    if (tid < 6 && cnt > 0) {
        // same-address atomics serialise in L2: peek first, only CTAs that actually move an extreme issue one
        const int32_t cur = *reinterpret_cast<volatile int32_t*>(L.minmax + tid), mine = s_mm[tid];
        if (tid & 1) { if (mine > cur) atomicMax(L.minmax + tid, mine); } else { if (mine < cur) atomicMin(L.minmax + tid, mine); }
    }
    cta_image_out(L.out + (dst0 - phase), s_img, phase, phase + cnt * kLasRec, tid, kLasThreads);   // TMA bulk store of the aligned body
    if (fl != 0 && L.status != nullptr) atomicOr(L.status, fl);
}

// LAS 1.2 public header block (227 bytes), written after the records so the extremes are final
__global__ void k_las_header(const __grid_constant__ LasParams L) {
    __shared__ uint8_t h[kLasHeader];
    const int tid = threadIdx.x;
    for (int i = tid; i < kLasHeader; i += blockDim.x) h[i] = 0;
    __syncthreads();
    if (tid == 0) {
        h[0] = 'L'; h[1] = 'A'; h[2] = 'S'; h[3] = 'F';
        h[24] = 1; h[25] = 2;                                                       // version 1.2
        const char sys[] = "OTHER"; for (int i = 0; i < 5; ++i) h[26 + i] = (uint8_t)sys[i];
        const char gen[] = "livox_mc_b200"; for (int i = 0; i < 13; ++i) h[58 + i] = (uint8_t)gen[i];
        put_le(h + 90, L.day, 2); put_le(h + 92, L.year, 2);
        put_le(h + 94, kLasHeader, 2);                                              // header size
        put_le(h + 96, kLasHeader, 4);                                              // offset to point data
        h[104] = 3;                                                                 // point data format
        put_le(h + 105, kLasRec, 2);
        put_le(h + 107, (uint64_t)L.n, 4);                                          // number of point records
        // number of points by return (5 x u32) stays 0: the reference never sets return numbers
        for (int c = 0; c < 3; ++c) {
            put_le(h + 131 + 8 * c, (uint64_t)__double_as_longlong(L.scale[c]), 8);
            put_le(h + 155 + 8 * c, (uint64_t)__double_as_longlong(L.off[c]), 8);
            const bool any = L.n > 0;
            const double vmax = any ? __fma_rn((double)L.minmax[2 * c + 1], L.scale[c], L.off[c]) : 0.0;   // X * scale + offset
            const double vmin = any ? __fma_rn((double)L.minmax[2 * c], L.scale[c], L.off[c]) : 0.0;
            put_le(h + 179 + 16 * c, (uint64_t)__double_as_longlong(vmax), 8);
            put_le(h + 187 + 16 * c, (uint64_t)__double_as_longlong(vmin), 8);
        }
    }
    __syncthreads();
    for (int i = tid; i < kLasHeader; i += blockDim.x) L.out[i] = h[i];
}

__global__ void k_las_init(int32_t* mm) { if (threadIdx.x < 6) mm[threadIdx.x] = (threadIdx.x & 1) ? INT32_MIN : INT32_MAX; }

// parts: 1 = records of points [p_begin, n) + min / max reduction into L.minmax, 2 = header from L.minmax, 3 = both
cudaError_t launch_las_pf3(bool f64, const LasParams& L, int parts, cudaStream_t st) {
    const int64_t tiles = (L.n - L.p_begin + kLasTile - 1) / kLasTile;
    if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
    if (parts & 1) {
        k_las_init<<<1, 32, 0, st>>>(L.minmax);
        if (tiles > 0) {
            if (f64) k_las_records<true><<<(unsigned)tiles, kLasThreads, 0, st>>>(L);
            else     k_las_records<false><<<(unsigned)tiles, kLasThreads, 0, st>>>(L);
        }
    }
    if (parts & 2) k_las_header<<<1, 256, 0, st>>>(L);
    return cudaGetLastError();
}

}  // namespace lmc
