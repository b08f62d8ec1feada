// lmc_lvx2.cu -- (second half) the LVX containers of the complete simulator's
// LivoxLVXWriter (CS:235-374) built on the device, so the file is one write() of a device buffer:
//
//   LMC_LVXCS_LVX2    _write_lvx2 (CS:269-287; _write_lvx3 emits the same bytes, CS:289-293):
//                     prefix = 24-B file header + 64-B private header (CS:323-341), then per frame a
//                     24-B header {u32 frame idx, u64 ts, u32 n, 8 x 0} (CS:348-352), ONE 21-byte package
//                     header per frame (CS:354-363 -- the comment there says 22) and n unpadded 14-B
//                     records {int(x*1000) x3 '<iii' (trunc, no clip), u8 intensity, u8 tag} (CS:365-374)
//   LMC_LVXCS_LEGACY  _write_lvx_legacy (CS:256-267, 295-321): prefix = 28-B header + 32-B device block,
//                     per frame {u64 ts, u32 n} and 14-B records {f32 x y z '<fff', u8 intensity, u8 tag}
//
// Both layouts are closed form: frame f starts at prefix_len + H*f + 14*frame_off[f] (H = 45 | 12), so no
// offset table is needed.  The prefix depends only on DeviceInfo / the frame count; the host builds
// those <= 96 bytes (lvx.py) and they travel in the launch parameters.
//
// One CTA = up to kCsChunk consecutive points of one frame (grid.x = frame, grid.y = chunk; chunk 0 also
// owns the frame's headers, so empty frames still get theirs).  The records start at arbitrary byte
// parity (45*f), so the CTA assembles its contiguous byte range in shared memory at the destination's
// 16-byte phase and copies it out as one TMA bulk store (byte stores on the ragged ends): every output
// byte is written exactly once.
#include "lmc_device.cuh"

namespace lmc {

constexpr int kCsChunk   = 1024;                    // points per CTA
constexpr int kCsThreads = 256;
constexpr int kCsHdrMax  = 45;
constexpr int kCsImg     = ((16 + kCsHdrMax + kCsChunk * 14 + 16 + 15) / 16) * 16;

struct LvxCsParams {
    const void*     pts;
    const uint8_t*  tag;
    const int64_t*  frame_off;
    const uint64_t* frame_ts;
    uint8_t*        out;
    uint32_t*       status;
    int32_t         n_frames, format, prefix_len;
    uint8_t         prefix[96];
};

__device__ __forceinline__ void put_u64(uint8_t* p, uint64_t v) {
#pragma unroll
    for (int k = 0; k < 8; ++k) p[k] = (uint8_t)(v >> (8 * k));
}
__device__ __forceinline__ void put_u32(uint8_t* p, uint32_t v) {
#pragma unroll
    for (int k = 0; k < 4; ++k) p[k] = (uint8_t)(v >> (8 * k));
}

// 14 bytes {x, y, z, rt(16 bit)} at an arbitrary byte address: 16-bit stores on the even part
__device__ __forceinline__ void put_record(uint8_t* p, uint32_t x, uint32_t y, uint32_t z, uint32_t rt) {
    if ((reinterpret_cast<uintptr_t>(p) & 1) == 0) {
        uint16_t* h = reinterpret_cast<uint16_t*>(p);
        h[0] = (uint16_t)x; h[1] = (uint16_t)(x >> 16); h[2] = (uint16_t)y; h[3] = (uint16_t)(y >> 16);
        h[4] = (uint16_t)z; h[5] = (uint16_t)(z >> 16); h[6] = (uint16_t)rt;
    } else {
        p[0] = (uint8_t)x;
        uint16_t* h = reinterpret_cast<uint16_t*>(p + 1);
        h[0] = (uint16_t)(x >> 8);
        h[1] = (uint16_t)((x >> 24) | (y << 8));
        h[2] = (uint16_t)(y >> 8);
        h[3] = (uint16_t)((y >> 24) | (z << 8));
        h[4] = (uint16_t)(z >> 8);
        h[5] = (uint16_t)((z >> 24) | (rt << 8));
        p[13] = (uint8_t)(rt >> 8);
    }
}

// struct.pack('<f', v): round to nearest f32; a finite value that rounds to inf raises OverflowError
__device__ __forceinline__ uint32_t q_f32_pack(double v, uint32_t& fl) {
    const float f = __double2float_rn(v);
    if (isinf(f) && !isinf(v)) fl |= LMC_FLAG_OVERFLOW;
    return __float_as_uint(f);
}

template <bool F64>
__global__ void __launch_bounds__(kCsThreads) k_lvx_cs(const __grid_constant__ LvxCsParams P) {
    __shared__ __align__(16) uint8_t s_img[kCsImg];
    const int f = blockIdx.x, chunk = blockIdx.y, tid = threadIdx.x;
    const int64_t p0 = P.frame_off[f], n = P.frame_off[f + 1] - p0;
    const int64_t first = (int64_t)chunk * kCsChunk;
    if (f == 0 && chunk == 0 && tid < P.prefix_len) P.out[tid] = P.prefix[tid];
    if (chunk != 0 && first >= n) return;                                             // empty frames still own their headers
    const bool lvx2 = P.format == LMC_LVXCS_LVX2;
    const int H = lvx2 ? 45 : 12;
    const int hdr = chunk == 0 ? H : 0;
    const int npts = (int)min((int64_t)kCsChunk, n - first > 0 ? n - first : 0);
    const int64_t dst0 = (int64_t)P.prefix_len + (int64_t)H * f + 14 * p0 + (chunk == 0 ? 0 : H + 14 * first);
    const int nbytes = hdr + 14 * npts;
    const int phase = (int)(dst0 & 15);
    uint8_t* img = s_img + phase;                                                     // img[i] <-> out[dst0 + i]

    if (chunk == 0 && tid == 0) {
        const uint64_t ts = P.frame_ts[f];
        if (lvx2) {
            put_u32(img, (uint32_t)f); put_u64(img + 4, ts); put_u32(img + 12, (uint32_t)n); put_u64(img + 16, 0);   // CS:348-352
            uint8_t* h = img + 24;                                                                                   // CS:354-363
            h[0] = 5; h[1] = 0; h[2] = 1; h[3] = 0; put_u32(h + 4, 0); h[8] = 1; h[9] = 2; h[10] = 0; h[11] = 0; h[12] = 0;
            put_u64(h + 13, ts);
        } else {
            put_u64(img, ts); put_u32(img + 8, (uint32_t)n);                                                         // CS:313-315
        }
    }
    uint32_t fl = 0;
    constexpr int K = kCsChunk / kCsThreads;
    RawRow<F64> raw[K];
    uint32_t tg[K];
    load_rows_strided<F64, K, kCsThreads>(P.pts, p0 + first, tid, npts, raw);
#pragma unroll
    for (int k = 0; k < K; ++k) { const int j = tid + k * kCsThreads; tg[k] = (P.tag && j < npts) ? (uint32_t)__ldg(P.tag + (p0 + first + j)) : 0u; }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int j = tid + k * kCsThreads;
        if (j >= npts) break;
        const Pt p = raw[k].pt();
        uint32_t x, y, z;
        if (lvx2) { x = (uint32_t)q_mm_noclip(p.x, fl); y = (uint32_t)q_mm_noclip(p.y, fl); z = (uint32_t)q_mm_noclip(p.z, fl); }   // CS:368-372
        else      { x = q_f32_pack(p.x, fl); y = q_f32_pack(p.y, fl); z = q_f32_pack(p.z, fl); }                                     // CS:319
        put_record(img + hdr + 14 * j, x, y, z, q_u8_copy(p.w, fl) | (tg[k] << 8));                                                  // CS:373-374 / 320-321
    }
    cta_image_out(P.out + (dst0 - phase), s_img, phase, phase + nbytes, tid, kCsThreads);      // TMA bulk store of the aligned body
    if (fl != 0 && P.status != nullptr) atomicOr(P.status, fl);
}

cudaError_t launch_lvx_cs(bool f64, const void* pts, const uint8_t* tag, const int64_t* frame_off, const uint64_t* frame_ts,
                          const uint8_t* prefix_host, int32_t prefix_len, int32_t format, uint8_t* out, int32_t n_frames,
                          int64_t max_frame_points, uint32_t* status, cudaStream_t st) {
    LvxCsParams P;
    P.pts = pts; P.tag = tag; P.frame_off = frame_off; P.frame_ts = frame_ts; P.out = out; P.status = status;
    P.n_frames = n_frames; P.format = format; P.prefix_len = prefix_len;
    for (int i = 0; i < 96; ++i) P.prefix[i] = i < prefix_len ? prefix_host[i] : 0;
    int64_t chunks = (max_frame_points + kCsChunk - 1) / kCsChunk;
    if (chunks < 1) chunks = 1;
    if (chunks > 65535) return cudaErrorInvalidValue;                                 // > 67 M points in one frame
    if (n_frames <= 0) {                                                              // header-only file
        if (prefix_len > 0) return cudaMemcpyAsync(out, prefix_host, (size_t)prefix_len, cudaMemcpyHostToDevice, st);
        return cudaSuccess;
    }
    dim3 grid((unsigned)n_frames, (unsigned)chunks);
    if (f64) k_lvx_cs<true><<<grid, kCsThreads, 0, st>>>(P);
    else     k_lvx_cs<false><<<grid, kCsThreads, 0, st>>>(P);
    return cudaGetLastError();
}

}  // namespace lmc
