// lmc_pcd.cu --  ASCII PCD point data on the device: the per-point line of
// LiDARMotionSimulator.save_pcd (LMC:946-947)
//     f"{x:.6f} {y:.6f} {z:.6f} {intensity:.6f}\n"
// for every row of an (N,4) array, byte-identical to CPython's formatting (correctly rounded,
// round-half-even on the exact binary value -- the same rule as C printf).
//
// "%.6f" of a double exactly:  |v| = m * 2^e (m < 2^53)  ->  Q = round_half_even(m * 10^6 * 2^e)
// with the 73-bit product m * 10^6 held in unsigned __int128; then Q / 10^6 and Q % 10^6 are the
// integer and fractional digits.  Values with |v| >= 9.2e12 (Q would not fit 64 bits) are printed as
// "inf"-free saturated text and flagged LMC_FLAG_OVERFLOW -- no point cloud coordinate gets there.
// (The generic row writer below -- k_text_*, the complete simulator's exports CS:1643-1716 -- splits
// integer and fraction instead and covers |v| < 2^64 with 0..9 decimals per column.)
//
// Lines have variable length, so the text is produced in three launches:
//   k_pcd_len    per tile of 256 points: total text bytes of the tile.  The length of "%.6f" needs no digits: it is
//                sign + 7 + the digit count of the ROUNDED integer part, and |v| rounds up to 10^k exactly when
//                |v| >= 10^k - 5e-7 -- a real number no double equals, so comparing with the first double above it
//                (kDigitT) is exact.  Four compares per number; HBM-bound.
//   k_pcd_scan   one CTA: exclusive prefix sum of the tile sizes -> byte offset of every tile
//   k_pcd_write  per tile: format, lay the lines out in shared memory (block scan of the line lengths) and copy the
//                tile's contiguous byte range out as one TMA bulk store.  Numbers below 10^4 (every LiDAR coordinate)
//                take a branch-free path: digit pairs by arithmetic, characters packed into words in registers and
//                streamed into the image as aligned 32-bit stores (WordStream below); anything else (>= 10^4, nan,
//                inf) goes through the general formatter byte by byte.
#include "lmc_device.cuh"

namespace lmc {

constexpr int kPcdTile    = 256;                    // points (= threads) per tile
constexpr int kNumMax     = 1 + 13 + 1 + 6;         // sign + 13 integer digits + '.' + 6 decimals
constexpr int kLineMax    = 4 * kNumMax + 4;        // 3 spaces + newline
constexpr int kPcdImg     = ((kPcdTile * kLineMax + 32 + 15) / 16) * 16;

// One formatted number: integer part, 6 fractional digits as an integer, text length, special-value kind
struct Num { uint64_t ip; uint32_t fq; uint32_t len; uint32_t kind; bool neg; };   // kind 0 finite, 1 nan, 2 inf

// Exact Q = round_half_even(|v| * 1e6) through the 73-bit integer product (any magnitude below 9.2e12)
__device__ __noinline__ uint64_t fmt_q_wide(uint64_t m, int e, uint32_t& fl) {
    uint64_t q;
    if (e >= 0) {                                    // integer-valued, >= 2^52: beyond the supported magnitude
        fl |= LMC_FLAG_OVERFLOW; q = 9199999999999999999ull;
    } else {
        const int s = -e;
        const unsigned __int128 P = (unsigned __int128)m * 1000000u;
        if (s >= 127) q = 0;                         // far below half of the last place
        else {
            const unsigned __int128 Qw = P >> s;
            if (Qw > (unsigned __int128)9199999999999999999ull) { fl |= LMC_FLAG_OVERFLOW; q = 9199999999999999999ull; }
            else {
                q = (uint64_t)Qw;
                const unsigned __int128 rem = P - (Qw << s), half = (unsigned __int128)1 << (s - 1);
                if (rem > half || (rem == half && (q & 1ull))) q += 1;
            }
        }
    }
    return q;
}

// "%.6f" of v.  Fast path for |v| < 2^43 (every coordinate a LiDAR produces), all in exact FP64 steps:
//   ip = trunc(|v|), fr = |v| - ip (exact), hi + lo = fr * 1e6 exactly (FMA error term),
//   q = floor(hi); the tie test t = (hi - q) - 0.5 is exact whenever |t| is small, and |lo| <= ulp(hi)/2
//   can only decide when t == 0  ->  round-half-even on the exact binary value, as printf does.
__device__ __forceinline__ Num fmt_prepare(double v, uint32_t& fl) {
    Num t;
    const uint64_t bits = (uint64_t)__double_as_longlong(v);
    t.neg = bits >> 63;
    t.kind = 0;
    const double a = fabs(v);
    if (a < 8796093022208.0) {                       // 2^43
        uint64_t ip = __double2ull_rz(a);
        const double fr = __dsub_rn(a, __ull2double_rn(ip));
        const double hi = __dmul_rn(fr, 1.0e6), lo = __fma_rn(fr, 1.0e6, -hi);
        uint32_t q = __double2uint_rz(hi);
        const double d = __dsub_rn(__dsub_rn(hi, __uint2double_rn(q)), 0.5);
        if (d > 0.0 || (d == 0.0 && (lo > 0.0 || (lo == 0.0 && (q & 1u))))) q += 1;
        if (q == 1000000u) { q = 0; ip += 1; }
        t.ip = ip; t.fq = q;
    } else {
        const uint32_t ex = (uint32_t)(bits >> 52) & 0x7ffu;
        const uint64_t frac = bits & 0xfffffffffffffull;
        t.ip = 0; t.fq = 0;
        if (ex == 0x7ff) { t.kind = frac ? 1u : 2u; t.len = frac ? 3u : (t.neg ? 4u : 3u); return t; }   // "nan" | "inf" / "-inf"
        const uint64_t Q = fmt_q_wide(frac | (1ull << 52), (int)ex - 1075, fl);
        t.ip = Q / 1000000ull; t.fq = (uint32_t)(Q - t.ip * 1000000ull);
    }
    uint32_t nd;
    if (t.ip < 1000000ull) {
        const uint32_t x = (uint32_t)t.ip;
        nd = 1u + (x >= 10u) + (x >= 100u) + (x >= 1000u) + (x >= 10000u) + (x >= 100000u);
    } else {
        nd = 7;
        for (uint64_t x = t.ip / 1000000ull; x >= 10; x /= 10) ++nd;
    }
    t.len = (t.neg ? 1u : 0u) + nd + 7u;
    return t;
}

__device__ __forceinline__ int fmt_write(uint8_t* dst, const Num& t) {
    if (t.kind) {
        int o = 0;
        if (t.kind == 2 && t.neg) dst[o++] = '-';
        if (t.kind == 1) { dst[o] = 'n'; dst[o + 1] = 'a'; dst[o + 2] = 'n'; }
        else             { dst[o] = 'i'; dst[o + 1] = 'n'; dst[o + 2] = 'f'; }
        return o + 3;
    }
    int o = (int)t.len;
    uint32_t fp = t.fq;
#pragma unroll
    for (int k = 0; k < 6; ++k) { dst[--o] = (uint8_t)('0' + fp % 10u); fp /= 10u; }
    dst[--o] = '.';
    if (t.ip < 1000000000ull) {                      // 32-bit digit loop
        uint32_t x = (uint32_t)t.ip;
        do { dst[--o] = (uint8_t)('0' + x % 10u); x /= 10u; } while (x);
    } else {
        uint64_t x = t.ip;
        do { dst[--o] = (uint8_t)('0' + (uint32_t)(x % 10ull)); x /= 10ull; } while (x);
    }
    if (t.neg) dst[--o] = '-';
    return (int)t.len;
}

// ---- fast path: |v| below 10^4 after rounding (at most 4 integer digits) ------------------------------------------
// kDigitT[k-1] = the smallest double >= 10^k - 5e-7 (tests/test_host.py re-derives them with exact fractions); for a
// float-valued |v| the same thresholds are 10, 100, 1000, 10000 (the float just below 10^k prints 9...9.99....).
__device__ __constant__ double kDigitT[4] = { 0x1.3ffffef390860p+3, 0x1.8fffffde7210cp+6, 0x1.f3fffffbce422p+9, 0x1.387fffffbce43p+13 };
constexpr uint32_t kSlowBit = 0x80000000u;                          // tags a length: the value needs the general formatter

__device__ __noinline__ uint32_t fmt_len_slow(double v, uint32_t& fl) { return fmt_prepare(v, fl).len; }
__device__ __noinline__ int fmt_write_slow(uint8_t* dst, double v, uint32_t& fl) { const Num t = fmt_prepare(v, fl); return fmt_write(dst, t); }

// text length of "%.6f" % v, without any digit: four compares
__device__ __forceinline__ uint32_t fmt_len(double v, uint32_t& fl) {
    const double a = fabs(v);
    if (!(a < kDigitT[3])) return fmt_len_slow(v, fl);              // >= 10^4, inf, nan
    return ((uint32_t)__double2hiint(v) >> 31) + 8u + (a >= kDigitT[0]) + (a >= kDigitT[1]) + (a >= kDigitT[2]);
}
__device__ __forceinline__ uint32_t fmt_len(float v, uint32_t& fl) {
    const float a = fabsf(v);
    if (!(a < 10000.0f)) return fmt_len_slow((double)v, fl);
    return (__float_as_uint(v) >> 31) + 8u + (a >= 10.0f) + (a >= 100.0f) + (a >= 1000.0f);
}

// One number ready to print: rounded integer part, the 6 fractional digits as an integer, tagged text length
struct Prep { uint32_t ip, q, len; };

__device__ __forceinline__ void prep_finish(Prep& p, uint32_t ip, uint32_t q, uint32_t neg) {
    const uint32_t carry = q == 1000000u ? 1u : 0u;                  // 0.9999995 -> 1.000000
    p.q = carry ? 0u : q;
    p.ip = ip + carry;
    p.len = neg + 8u + (p.ip >= 10u) + (p.ip >= 100u) + (p.ip >= 1000u);
}
// float4 layout: |v| - trunc(|v|) is exact in f32 and its product with 10^6 = 2^6 * 15625 is exact in f64 (24 + 14 bits),
// so ONE round-to-nearest-even conversion is printf's rounding of the exact binary value -- no tie handling at all
__device__ __forceinline__ Prep fmt_prep(float v, uint32_t& fl) {
    Prep p;
    const float a = fabsf(v);
    if (!(a < 10000.0f)) { p.ip = p.q = 0; p.len = fmt_len_slow((double)v, fl) | kSlowBit; return p; }
    const uint32_t ip = __float2uint_rz(a);
    const float fr = __fsub_rn(a, __uint2float_rn(ip));
    prep_finish(p, ip, __double2uint_rn(__dmul_rn((double)fr, 1.0e6)), __float_as_uint(v) >> 31);
    return p;
}
// f64: hi + lo = fr * 10^6 exactly (FMA error term).  RNE(hi) is the answer unless hi sits exactly on a tie k + 0.5 (the only
// place where the sign of lo can change the decision, since k + 0.5 is itself a double and rounding is monotonic)
__device__ __forceinline__ Prep fmt_prep(double v, uint32_t& fl) {
    Prep p;
    const double a = fabs(v);
    if (!(a < kDigitT[3])) { p.ip = p.q = 0; p.len = fmt_len_slow(v, fl) | kSlowBit; return p; }
    const uint32_t ip = __double2uint_rz(a);
    const double fr = __dsub_rn(a, __uint2double_rn(ip));
    const double hi = __dmul_rn(fr, 1.0e6);
    uint32_t q = __double2uint_rn(hi);
    if (fabs(__dsub_rn(hi, __uint2double_rn(q))) == 0.5) {
        const double lo = __fma_rn(fr, 1.0e6, -hi);
        const uint32_t dn = __double2uint_rz(hi);
        if (lo > 0.0) q = dn + 1u; else if (lo < 0.0) q = dn;
    }
    prep_finish(p, ip, q, (uint32_t)__double2hiint(v) >> 31);
    return p;
}

// ---- text assembly: a per-thread word stream into the tile image ------------------------------------------------------
// Byte stores into shared memory were the bottleneck of the first version of this kernel (52 STS.U8 + 20 table LDS per line,
// ~170 shared-memory wavefronts per warp: the LSU data pipe was 86 % busy).  Now every thread streams its line as aligned
// 32-bit words: characters are packed in registers (digit pairs by arithmetic, no table), a 1..4-byte piece is appended to
// a pending word with two funnel shifts, and the fixed 8-byte tail ".dddddd<sep>" of every number leaves as two whole words.
// Only the first and the last word of a line are shared with the neighbouring lines: those go out with atomicOr into the
// zero-initialised image (OR commutes, so the result does not depend on the order of the threads).
__device__ __forceinline__ uint32_t digit_pair(uint32_t n) {          // n < 100 -> '0' + n / 10 | ('0' + n % 10) << 8
    const uint32_t t = (n * 205u) >> 11;
    return 0x3030u + t + ((n - 10u * t) << 8);
}
struct WordStream {
    uint32_t* img;        // tile image as words (zero-initialised)
    uint32_t  wp;         // index of the pending word
    uint32_t  n;          // bytes already in the pending word (0..3); bytes [0, n) of a line's FIRST word belong to the previous line
    uint32_t  a0;         // the pending word
    __device__ __forceinline__ void start(uint32_t* image, uint32_t byte_pos) { img = image; wp = byte_pos >> 2; n = byte_pos & 3u; a0 = 0; }
    // append the k (0..4) low bytes of chunk (its other bytes must be 0).  FIRST: the word being completed may be the line's first
    template <bool FIRST>
    __device__ __forceinline__ void put(uint32_t chunk, uint32_t k) {
        const uint32_t sh = 8u * n;
        a0 |= chunk << sh;
        const uint32_t hi = __funnelshift_l(chunk, 0u, sh);           // the bytes that did not fit (0 when n == 0)
        n += k;
        if (n >= 4u) {
            if (FIRST) atomicOr(img + wp, a0); else img[wp] = a0;
            wp += 1u; a0 = hi; n -= 4u;
        }
    }
    // append 8 bytes (two whole words leave, the number of pending bytes is unchanged)
    template <bool FIRST>
    __device__ __forceinline__ void put8(uint32_t lo, uint32_t hi) {
        const uint32_t sh = 8u * n;
        const uint32_t w0 = a0 | (lo << sh);
        if (FIRST) atomicOr(img + wp, w0); else img[wp] = w0;
        img[wp + 1u] = __funnelshift_l(lo, hi, sh);
        a0 = __funnelshift_l(hi, 0u, sh);
        wp += 2u;
    }
    __device__ __forceinline__ void finish() { if (n) atomicOr(img + wp, a0); }   // the line's last word is the next line's first
};

// one prepared number + separator into the stream
template <bool FIRST>
__device__ __forceinline__ void fmt_stream(WordStream& ws, const Prep& p, double v, uint32_t sep, uint32_t& fl) {
    if (p.len & kSlowBit) {                                          // general formatter (rare): bytes through the same stream
        uint8_t tmp[kNumMax + 3];
        const int len = fmt_write_slow(tmp, v, fl);
        for (int k = 0; k < len; ++k) ws.put<FIRST>(tmp[k], 1u);
        ws.put<FIRST>(sep, 1u);
        return;
    }
    const uint32_t neg = (uint32_t)__double2hiint(v) >> 31;
    const uint32_t nd = p.len - 7u - neg;
    const uint32_t q1 = p.q / 10000u, r = p.q - q1 * 10000u, q2 = (r * 5243u) >> 19, q3 = r - q2 * 100u;   // r < 10^4: r / 100 exactly
    const uint32_t p1 = digit_pair(q1), p2 = digit_pair(q2), p3 = digit_pair(q3);
    const uint32_t i1 = (p.ip * 5243u) >> 19, i0 = p.ip - i1 * 100u;
    const uint32_t digits = (digit_pair(i1) | (digit_pair(i0) << 16)) >> (8u * (4u - nd));   // the nd integer digits, first digit in byte 0
    ws.put<FIRST>(neg ? 0x2du : 0u, neg);
    ws.put<FIRST>(digits, nd);
    ws.put8<FIRST>(0x2eu | (p1 << 8) | (p2 << 24), (p2 >> 8) | (p3 << 8) | (sep << 24));                  // ".dddddd" + separator
}

template <bool F64>
__device__ __forceinline__ void load_row(const void* pts, int64_t i, double (&v)[4]) {
    if constexpr (F64) { const double* s = reinterpret_cast<const double*>(pts) + 4 * i; ldg256(s, v[0], v[1], v[2], v[3]); }
    else { const float4 f = __ldg(reinterpret_cast<const float4*>(pts) + i); v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w; }
}
// the row in its own type (float4 rows are formatted from the floats: no conversion, f32 compares)
template <bool F64> struct RowT { using T = double; };
template <> struct RowT<false> { using T = float; };
template <bool F64>
__device__ __forceinline__ void load_row_native(const void* pts, int64_t i, typename RowT<F64>::T (&v)[4]) {
    if constexpr (F64) { const double* s = reinterpret_cast<const double*>(pts) + 4 * i; ldg256(s, v[0], v[1], v[2], v[3]); }
    else { const float4 f = __ldg(reinterpret_cast<const float4*>(pts) + i); v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w; }
}

__device__ __forceinline__ uint32_t block_scan_excl(uint32_t x, uint32_t* s_warp, uint32_t& total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < kPcdTile / 32; ++k) { const uint32_t t = s_warp[k]; if (k < w) base += t; tot += t; }
    total = tot;
    return base + inc - x;
}

template <bool F64>
__global__ void __launch_bounds__(kPcdTile) k_pcd_len(const void* __restrict__ pts, int64_t n, int64_t* __restrict__ tile_off) {
    __shared__ uint32_t s_warp[kPcdTile / 32];
    const int64_t i = (int64_t)blockIdx.x * kPcdTile + threadIdx.x;
    uint32_t len = 0, fl = 0;
    if (i < n) {
        typename RowT<F64>::T v[4];
        load_row_native<F64>(pts, i, v);
        len = 4;
#pragma unroll
        for (int c = 0; c < 4; ++c) len += fmt_len(v[c], fl);
    }
    uint32_t total;
    block_scan_excl(len, s_warp, total);
    if (threadIdx.x == 0) tile_off[blockIdx.x + 1] = total;       // sizes now, offsets after k_pcd_scan
}

// one CTA: tile_off[0] = 0; tile_off[t+1] = sum of sizes[0..t]   (in place)
__global__ void __launch_bounds__(1024) k_pcd_scan(int64_t* __restrict__ tile_off, int64_t n_tiles) {
    __shared__ int64_t s_part[1024];
    const int t = threadIdx.x;
    const int64_t per = (n_tiles + 1023) / 1024;
    const int64_t b = t * per, e = b + per < n_tiles ? b + per : n_tiles;
    int64_t sum = 0;
    for (int64_t k = b; k < e; ++k) sum += tile_off[k + 1];
    s_part[t] = sum;
    __syncthreads();
    if (t == 0) { int64_t acc = 0; for (int k = 0; k < 1024; ++k) { const int64_t v = s_part[k]; s_part[k] = acc; acc += v; } tile_off[0] = 0; }
    __syncthreads();
    int64_t acc = s_part[t];
    for (int64_t k = b; k < e; ++k) { acc += tile_off[k + 1]; tile_off[k + 1] = acc; }
}

template <bool F64>
__global__ void __launch_bounds__(kPcdTile) k_pcd_write(const void* __restrict__ pts, int64_t n, const int64_t* __restrict__ tile_off,
                                                        uint8_t* __restrict__ out, uint32_t* __restrict__ status) {
    __shared__ __align__(16) uint8_t s_img[kPcdImg];
    __shared__ uint32_t s_warp[kPcdTile / 32];
    const int tid = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * kPcdTile + tid;
    const int64_t dst0 = tile_off[blockIdx.x];
    const int phase = (int)(dst0 & 15);
    // zero the part of the image this tile can reach (the first / last word of every line is OR-ed in)
    const int img_words = (int)((phase + (tile_off[blockIdx.x + 1] - dst0) + 3 + 15) / 16) * 4;
    for (int k = tid * 4; k < img_words; k += kPcdTile * 4) *reinterpret_cast<uint4*>(s_img + 4 * k) = make_uint4(0, 0, 0, 0);
    typename RowT<F64>::T v[4] = { 0, 0, 0, 0 };
    Prep pr[4];
    uint32_t len = 0, fl = 0;
    if (i < n) {
        load_row_native<F64>(pts, i, v);
        len = 4;
#pragma unroll
        for (int c = 0; c < 4; ++c) { pr[c] = fmt_prep(v[c], fl); len += pr[c].len & ~kSlowBit; }
    }
    uint32_t total;
    const uint32_t off = block_scan_excl(len, s_warp, total);       // (its barrier also orders the zero fill before the ORs)
    if (i < n) {
        WordStream ws;
        ws.start(reinterpret_cast<uint32_t*>(s_img), (uint32_t)phase + off);
        fmt_stream<true>(ws, pr[0], (double)v[0], ' ', fl);
        fmt_stream<false>(ws, pr[1], (double)v[1], ' ', fl);
        fmt_stream<false>(ws, pr[2], (double)v[2], ' ', fl);
        fmt_stream<false>(ws, pr[3], (double)v[3], '\n', fl);
        ws.finish();
    }
    cta_image_out(out + (dst0 - phase), s_img, phase, phase + (int)total, tid, kPcdTile);        // TMA bulk store of the aligned body
    if (fl != 0 && status != nullptr) atomicOr(status, fl);
}

// Byte offset of the text of arbitrary rows (frame boundaries): one warp per query sums the line lengths of
// the rows of its tile that precede the row.  With these offsets ONE formatting pass over the frame-major
// buffer yields every per-frame file body of save_results (LMC:870-884) as a slice of the same text.
template <bool F64>
__global__ void __launch_bounds__(256) k_pcd_row_off(const void* __restrict__ pts, int64_t n, const int64_t* __restrict__ tile_off,
                                                     const int64_t* __restrict__ rows, int32_t n_rows, int64_t* __restrict__ byte_off) {
    const int lane = threadIdx.x & 31;
    const int64_t qi = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (qi >= n_rows) return;
    int64_t r = rows[qi];
    r = r < 0 ? 0 : (r > n ? n : r);
    const int64_t tile = r / kPcdTile, first = tile * kPcdTile;
    uint32_t sum = 0, fl = 0;
    for (int64_t i = first + lane; i < r; i += 32) {
        typename RowT<F64>::T v[4];
        load_row_native<F64>(pts, i, v);
        sum += 4;
#pragma unroll
        for (int c = 0; c < 4; ++c) sum += fmt_len(v[c], fl);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) byte_off[qi] = tile_off[tile] + sum;
}

// ---- generic delimited text rows: CS:1643-1716 (_export_pcd '%.6f %.6f %.6f %.0f %.0f', _export_xyz, _export_csv) ----
// "%.{d}f" with d = 0..9 per column and the full integer range a timestamp column needs (ns since the
// epoch ~ 1.7e18): the value is split exactly into its integer part I (< 2^64) and the fraction
// Q = round_half_even(frac * 10^d) (carry into I when Q reaches 10^d), all in exact FP64 steps.
constexpr int kTextCols   = 6;                      // LMC_TEXT_MAX_COLS

struct TextFmt { int32_t n_cols, row_stride; int32_t col[kTextCols]; int32_t dec[kTextCols]; uint8_t sep; };

__device__ __constant__ uint32_t kPow10[10] = { 1u, 10u, 100u, 1000u, 10000u, 100000u, 1000000u, 10000000u, 100000000u, 1000000000u };

struct TextNum { uint64_t ip; uint32_t fq; uint32_t len; uint8_t kind; bool neg; };   // kind 0 finite, 1 nan, 2 inf

__device__ __constant__ double kPow10d[10] = { 1.0, 10.0, 100.0, 1.0e3, 1.0e4, 1.0e5, 1.0e6, 1.0e7, 1.0e8, 1.0e9 };

// Same exact-FP64 scheme as fmt_prepare: ip = trunc(|v|) (exact below 2^64), fr = |v| - ip (exact),
// hi + lo = fr * 10^d exactly (10^d has <= 21 significant bits), q = floor(hi), tie test on (hi - q) - 0.5
// with lo as the tie breaker; ties go to the even last printed digit (ip's when d == 0).
__device__ __forceinline__ TextNum fmtg_prepare(double v, int d, uint32_t& fl) {
    TextNum t;
    const uint64_t bits = (uint64_t)__double_as_longlong(v);
    t.neg = bits >> 63;
    t.ip = 0; t.fq = 0; t.kind = 0;
    const double a = fabs(v);
    if (a < 18446744073709551616.0) {                // 2^64
        uint64_t ip = __double2ull_rz(a);
        const double fr = __dsub_rn(a, __ull2double_rn(ip));
        const double p10 = kPow10d[d];
        const double hi = __dmul_rn(fr, p10), lo = __fma_rn(fr, p10, -hi);
        uint32_t q = __double2uint_rz(hi);
        const double dd = __dsub_rn(__dsub_rn(hi, __uint2double_rn(q)), 0.5);
        const bool odd = d ? (q & 1u) : (uint32_t)(ip & 1ull);
        if (dd > 0.0 || (dd == 0.0 && (lo > 0.0 || (lo == 0.0 && odd)))) q += 1;
        if (q >= kPow10[d]) { q -= kPow10[d]; ip += 1; }
        t.ip = ip; t.fq = q;
    } else if (a != a) {
        t.kind = 1; t.len = 3u; return t;
    } else if (isinf(a)) {
        t.kind = 2; t.len = t.neg ? 4u : 3u; return t;
    } else {
        fl |= LMC_FLAG_OVERFLOW; t.ip = 0xffffffffffffffffull;
    }
    uint32_t nd;
    if (t.ip < 1000000ull) {
        const uint32_t x = (uint32_t)t.ip;
        nd = 1u + (x >= 10u) + (x >= 100u) + (x >= 1000u) + (x >= 10000u) + (x >= 100000u);
    } else {
        nd = 7;
        for (uint64_t x = t.ip / 1000000ull; x >= 10; x /= 10) ++nd;
    }
    t.len = (t.neg ? 1u : 0u) + nd + (d ? 1u + (uint32_t)d : 0u);
    return t;
}

__device__ __forceinline__ int fmtg_write(uint8_t* dst, int d, const TextNum& t) {
    if (t.kind) {
        int o = 0;
        if (t.kind == 2 && t.neg) dst[o++] = '-';
        if (t.kind == 1) { dst[o] = 'n'; dst[o + 1] = 'a'; dst[o + 2] = 'n'; }
        else             { dst[o] = 'i'; dst[o + 1] = 'n'; dst[o + 2] = 'f'; }
        return o + 3;
    }
    int o = (int)t.len;
    if (d) {
        uint32_t fp = t.fq;
        for (int k = 0; k < d; ++k) { dst[--o] = (uint8_t)('0' + fp % 10u); fp /= 10u; }
        dst[--o] = '.';
    }
    if (t.ip < 1000000000ull) {
        uint32_t x = (uint32_t)t.ip;
        do { dst[--o] = (uint8_t)('0' + x % 10u); x /= 10u; } while (x);
    } else {
        uint64_t ip = t.ip;
        do { dst[--o] = (uint8_t)('0' + (uint32_t)(ip % 10ull)); ip /= 10ull; } while (ip);
    }
    if (t.neg) dst[--o] = '-';
    return (int)t.len;
}

template <bool F64>
__device__ __forceinline__ double load_cell(const void* rows, int64_t idx) {
    if constexpr (F64) return __ldg(reinterpret_cast<const double*>(rows) + idx);
    else return (double)__ldg(reinterpret_cast<const float*>(rows) + idx);
}

template <bool F64>
__global__ void __launch_bounds__(kPcdTile) k_text_len(const void* __restrict__ rows, int64_t n, const __grid_constant__ TextFmt F, int64_t* __restrict__ tile_off) {
    __shared__ uint32_t s_warp[kPcdTile / 32];
    const int64_t i = (int64_t)blockIdx.x * kPcdTile + threadIdx.x;
    uint32_t len = 0, fl = 0;
    if (i < n) {
        len = (uint32_t)F.n_cols;                                    // separators + newline
#pragma unroll
        for (int c = 0; c < kTextCols; ++c) if (c < F.n_cols) len += fmtg_prepare(load_cell<F64>(rows, i * F.row_stride + F.col[c]), F.dec[c], fl).len;
    }
    uint32_t total;
    block_scan_excl(len, s_warp, total);
    if (threadIdx.x == 0) tile_off[blockIdx.x + 1] = total;
}

template <bool F64>
__global__ void __launch_bounds__(kPcdTile) k_text_write(const void* __restrict__ rows, int64_t n, const __grid_constant__ TextFmt F,
                                                         const int64_t* __restrict__ tile_off, uint8_t* __restrict__ out, uint32_t* __restrict__ status) {
    extern __shared__ __align__(16) uint8_t s_dyn[];                 // the tile's text image, sized by the host from the format
    __shared__ uint32_t s_warp[kPcdTile / 32];
    uint8_t* s_img = s_dyn;
    const int tid = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * kPcdTile + tid;
    const int64_t dst0 = tile_off[blockIdx.x];
    const int phase = (int)(dst0 & 15);
    TextNum t[kTextCols];
    uint32_t len = 0, fl = 0;
    if (i < n) {
        len = (uint32_t)F.n_cols;
#pragma unroll
        for (int c = 0; c < kTextCols; ++c) if (c < F.n_cols) {
            t[c] = fmtg_prepare(load_cell<F64>(rows, i * F.row_stride + F.col[c]), F.dec[c], fl);
            len += t[c].len;
        }
    }
    uint32_t total;
    const uint32_t off = block_scan_excl(len, s_warp, total);
    if (i < n) {
        uint8_t* d = s_img + phase + off;
#pragma unroll
        for (int c = 0; c < kTextCols; ++c) if (c < F.n_cols) { d += fmtg_write(d, F.dec[c], t[c]); *d++ = c == F.n_cols - 1 ? (uint8_t)'\n' : F.sep; }
    }
    cta_image_out(out + (dst0 - phase), s_img, phase, phase + (int)total, tid, kPcdTile);
    if (fl != 0 && status != nullptr) atomicOr(status, fl);
}

static TextFmt make_fmt(int32_t n_cols, int32_t row_stride, const int32_t* col, const int32_t* dec, uint8_t sep, int& img_bytes) {
    TextFmt F{};
    F.n_cols = n_cols; F.row_stride = row_stride; F.sep = sep;
    int row = n_cols;
    for (int c = 0; c < n_cols; ++c) { F.col[c] = col[c]; F.dec[c] = dec[c]; row += 1 + 20 + (dec[c] ? 1 + dec[c] : 0); }
    img_bytes = ((kPcdTile * row + 32 + 15) / 16) * 16;
    return F;
}

cudaError_t launch_text_size(bool f64, const void* rows, int64_t n, int32_t n_cols, int32_t row_stride, const int32_t* col, const int32_t* dec,
                             uint8_t sep, int64_t* tile_off, cudaStream_t st) {
    const int64_t tiles = (n + kPcdTile - 1) / kPcdTile;
    if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
    int img;
    const TextFmt F = make_fmt(n_cols, row_stride, col, dec, sep, img);
    if (tiles > 0) {
        if (f64) k_text_len<true><<<(unsigned)tiles, kPcdTile, 0, st>>>(rows, n, F, tile_off);
        else     k_text_len<false><<<(unsigned)tiles, kPcdTile, 0, st>>>(rows, n, F, tile_off);
    }
    k_pcd_scan<<<1, 1024, 0, st>>>(tile_off, tiles);
    return cudaGetLastError();
}

cudaError_t launch_text_write(bool f64, const void* rows, int64_t n, int32_t n_cols, int32_t row_stride, const int32_t* col, const int32_t* dec,
                              uint8_t sep, const int64_t* tile_off, uint8_t* out, uint32_t* status, cudaStream_t st) {
    const int64_t tiles = (n + kPcdTile - 1) / kPcdTile;
    if (tiles == 0) return cudaSuccess;
    int img;
    const TextFmt F = make_fmt(n_cols, row_stride, col, dec, sep, img);
    cudaError_t e = cudaSuccess;
    if (img > 48 * 1024)                                             // widest formats only (6 columns x 9 decimals = 49 KB)
        e = f64 ? cudaFuncSetAttribute(k_text_write<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, img)
                : cudaFuncSetAttribute(k_text_write<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, img);
    if (e != cudaSuccess) return e;
    if (f64) k_text_write<true><<<(unsigned)tiles, kPcdTile, img, st>>>(rows, n, F, tile_off, out, status);
    else     k_text_write<false><<<(unsigned)tiles, kPcdTile, img, st>>>(rows, n, F, tile_off, out, status);
    return cudaGetLastError();
}

cudaError_t launch_pcd_size(bool f64, const void* pts, int64_t n, int64_t* tile_off, cudaStream_t st) {
    const int64_t tiles = (n + kPcdTile - 1) / kPcdTile;
    if (tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
    if (tiles > 0) {
        if (f64) k_pcd_len<true><<<(unsigned)tiles, kPcdTile, 0, st>>>(pts, n, tile_off);
        else     k_pcd_len<false><<<(unsigned)tiles, kPcdTile, 0, st>>>(pts, n, tile_off);
    }
    k_pcd_scan<<<1, 1024, 0, st>>>(tile_off, tiles);
    return cudaGetLastError();
}

cudaError_t launch_pcd_row_off(bool f64, const void* pts, int64_t n, const int64_t* tile_off, const int64_t* rows, int32_t n_rows,
                               int64_t* byte_off, cudaStream_t st) {
    if (n_rows <= 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n_rows + 7) / 8);
    if (f64) k_pcd_row_off<true><<<grid, 256, 0, st>>>(pts, n, tile_off, rows, n_rows, byte_off);
    else     k_pcd_row_off<false><<<grid, 256, 0, st>>>(pts, n, tile_off, rows, n_rows, byte_off);
    return cudaGetLastError();
}

cudaError_t launch_pcd_write(bool f64, const void* pts, int64_t n, const int64_t* tile_off, uint8_t* out, uint32_t* status, cudaStream_t st) {
    const int64_t tiles = (n + kPcdTile - 1) / kPcdTile;
    if (tiles == 0) return cudaSuccess;
    if (f64) k_pcd_write<true><<<(unsigned)tiles, kPcdTile, 0, st>>>(pts, n, tile_off, out, status);
    else     k_pcd_write<false><<<(unsigned)tiles, kPcdTile, 0, st>>>(pts, n, tile_off, out, status);
    return cudaGetLastError();
}

}  // namespace lmc
