// lmc_lvx.cu --  the whole LVX v1.1 byte stream of LivoxLVXWriter.write_compatible_lvx
// (LMC:58-250) built on the device: quantise the RAW points (LMC:252-272) and lay out the container
// around them, so the file is one write() of a device-built buffer.
//
// Layout (closed form, LMC:126-133):  88-byte preamble | per frame: 24-byte header
//   {cur offset u64, next offset u64 (0 for the last frame), frame id u64}, then ceil(n/96) packages of
//   22-byte header + 96 x 14-byte records, the tail package zero-padded (LMC:246-250).
// frame_pos[f] (byte offset of frame f, frame_pos[F] = file size) is computed by the caller from the
// CSR offsets (livox_motion_compensation_sim_b200/lvx.py::frame_layout) and passed in.
//
// One CTA = up to kPk consecutive packages of one frame (grid.x = frame, grid.y = chunk).  The CTA
// assembles its contiguous byte range in shared memory at the destination's 16-byte phase and copies
// it out as one TMA bulk store (byte stores on the ragged ends): every output byte is written exactly
// once, nothing outside the range is touched.
#include "lmc_device.cuh"

namespace lmc {

constexpr int kPk        = 8;                       // packages per CTA
constexpr int kPkPoints  = 96;                      // LMC:48
constexpr int kPkBytes   = 22 + kPkPoints * 14;     // 1366
constexpr int kLvxThreads = 256;
constexpr int kImg       = ((16 + 24 + kPk * kPkBytes + 16 + 15) / 16) * 16;

__device__ __forceinline__ void put16(uint8_t* img, int off, uint32_t v) {            // off is even
    *reinterpret_cast<uint16_t*>(img + off) = (uint16_t)v;
}

template <bool F64>
__global__ void __launch_bounds__(kLvxThreads) k_lvx_v11(const void* __restrict__ pts, const int64_t* __restrict__ frame_off,
                                                         const int64_t* __restrict__ frame_pos, const double* __restrict__ frame_time,
                                                         const int64_t* __restrict__ frame_id, uint8_t* __restrict__ out,
                                                         int32_t n_frames, int32_t f_begin, uint32_t* __restrict__ status)
{
    // out is the (possibly virtual) address of file byte 0: a rank that builds only frames [f_begin, f_end) passes
    // shard_buffer - file_position_of_its_first_byte, so every store lands inside its own buffer
    __shared__ __align__(16) uint8_t s_img[kImg];
    const int f = f_begin + blockIdx.x, chunk = blockIdx.y, tid = threadIdx.x;
    const int64_t p0 = frame_off[f], n = frame_off[f + 1] - p0;
    const int64_t pkgs = (n + kPkPoints - 1) / kPkPoints;
    const int64_t pk0 = (int64_t)chunk * kPk;
    if (f == 0 && chunk == 0 && tid < 88) {
        // public header (LMC:85-101), private header (LMC:103-112), device info (LMC:147-172)
        uint8_t b = 0;
        const char sig[] = "livox_tech";
        if (tid < 10) b = (uint8_t)sig[tid];
        else if (tid == 16 || tid == 17) b = 1;                                       // version 1.1.0.0
        else if (tid >= 20 && tid < 24) b = (uint8_t)(0xAC0EA767u >> (8 * (tid - 20)));
        else if (tid == 24) b = 50;                                                   // frame duration ms
        else if (tid == 28) b = 1;                                                    // device count
        else if (tid >= 29 && tid < 44) { const char sn[] = "3GGDJ6K00200101"; b = (uint8_t)sn[tid - 29]; }
        else if (tid == 29 + 33) b = 1;                                               // device type
        out[tid] = b;
    }
    if (pk0 >= pkgs && !(chunk == 0)) return;                                         // nothing in this chunk (empty frames still own a header)
    const int npk = (int)min((int64_t)kPk, pkgs - pk0 > 0 ? pkgs - pk0 : 0);
    const int hdr = chunk == 0 ? 24 : 0;
    const int64_t dst0 = frame_pos[f] + (chunk == 0 ? 0 : 24 + pk0 * kPkBytes);
    const int nbytes = hdr + npk * kPkBytes;
    const int phase = (int)(reinterpret_cast<uintptr_t>(out + dst0) & 15);
    uint8_t* img = s_img + phase;                                                     // img[i] <-> out[dst0 + i]; the address is even

    // the points first: K independent loads per thread in flight under the zero fill and the header work
    const int64_t first = pk0 * kPkPoints;
    const int npts = (int)min((int64_t)npk * kPkPoints, n - first > 0 ? n - first : 0);
    constexpr int K = (kPk * kPkPoints + kLvxThreads - 1) / kLvxThreads;
    RawRow<F64> raw[K];
    load_rows_strided<F64, K, kLvxThreads>(pts, p0 + first, tid, npts, raw);

    for (int i = tid; i < kImg / 16; i += kLvxThreads) reinterpret_cast<uint4*>(s_img)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();

    if (chunk == 0 && tid < 3) {                                                      // frame header LMC:179-193
        const uint64_t v = tid == 0 ? (uint64_t)frame_pos[f] : tid == 1 ? (f + 1 < n_frames ? (uint64_t)frame_pos[f + 1] : 0ull) : (uint64_t)frame_id[f];
#pragma unroll
        for (int k = 0; k < 4; ++k) put16(img, 8 * tid + 2 * k, (uint32_t)(v >> (16 * k)));
    }
    if (tid < npk) {                                                                  // package headers LMC:206-237
        uint8_t* h = img + hdr + tid * kPkBytes;
        h[1] = 5; h[3] = 1; h[9] = 1; h[10] = 2;                                      // version, lidar id, ts type, data type
        const int64_t ts = __double2ll_rz(__dmul_rn(frame_time[f], 1e9));             // int(t * 1e9), LMC:177
#pragma unroll
        for (int k = 0; k < 4; ++k) put16(h, 14 + 2 * k, (uint32_t)((uint64_t)ts >> (16 * k)));
    }
    uint32_t fl = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = tid + k * kLvxThreads;
        if (j >= npts) break;
        const Pt p = raw[k].pt();
        const uint32_t x = (uint32_t)q_mm_clip(p.x, fl), y = (uint32_t)q_mm_clip(p.y, fl), z = (uint32_t)q_mm_clip(p.z, fl), r = q_refl(p.w, fl);
        const int off = hdr + (j / kPkPoints) * kPkBytes + 22 + (j % kPkPoints) * 14;
        put16(img, off, x); put16(img, off + 2, x >> 16);
        put16(img, off + 4, y); put16(img, off + 6, y >> 16);
        put16(img, off + 8, z); put16(img, off + 10, z >> 16);
        put16(img, off + 12, r);                                                      // reflectivity, tag = 0
    }
    // img[0, nbytes) -> out[dst0, dst0 + nbytes): one TMA bulk store for the aligned body, byte-wise ends
    cta_image_out(out + (dst0 - phase), s_img, phase, phase + nbytes, tid, kLvxThreads);
    if (fl != 0 && status != nullptr) atomicOr(status, fl);
}

cudaError_t launch_lvx_v11(bool f64, const void* pts, const int64_t* frame_off, const int64_t* frame_pos, const double* frame_time,
                           const int64_t* frame_id, uint8_t* out, int32_t n_frames, int32_t f_begin, int32_t f_end, int64_t max_frame_points,
                           uint32_t* status, cudaStream_t st) {
    if (n_frames <= 0 || f_end <= f_begin) return cudaSuccess;
    const int64_t max_pk = (max_frame_points + kPkPoints - 1) / kPkPoints;
    int64_t chunks = (max_pk + kPk - 1) / kPk;
    if (chunks < 1) chunks = 1;
    if (chunks > 65535) return cudaErrorInvalidValue;                                 // > 50 M points in one frame
    dim3 grid((unsigned)(f_end - f_begin), (unsigned)chunks);
    if (f64) k_lvx_v11<true><<<grid, kLvxThreads, 0, st>>>(pts, frame_off, frame_pos, frame_time, frame_id, out, n_frames, f_begin, status);
    else     k_lvx_v11<false><<<grid, kLvxThreads, 0, st>>>(pts, frame_off, frame_pos, frame_time, frame_id, out, n_frames, f_begin, status);
    return cudaGetLastError();
}

}  // namespace lmc
