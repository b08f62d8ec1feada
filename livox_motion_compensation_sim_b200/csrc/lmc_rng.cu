// lmc_rng.cu -- host-side replay of NumPy's legacy global generator for the scanner's range noise.
//
// The reference draws its per-frame noise with np.random.normal(0, lidar_range_noise, (n, 3)) from the seeded
// GLOBAL RandomState (LMC:767; seeded at LMC:288), so a run only reproduces the reference's raw scans if exactly
// that stream is consumed.  It is sequential by construction -- MT19937 words feed a polar rejection loop -- and
// NumPy spends ~13 ns per normal on it: 40 of the 51 ms of a whole C2a run on the device.  This file restates
// the published algorithm (numpy/random/src/mt19937/mt19937.h and src/legacy/legacy-distributions.c:
// legacy_gauss / legacy_double) and splits it: the sequential part (words -> accepted (x1, x2) pairs) stays on
// one thread, the expensive part (sqrt(-2 log(r2) / r2), the same libm calls NumPy makes) fans out over threads.
// Bit-identical to np.random.normal, including the generator state left behind (tests/test_host.py).
//
// No device code here; built into liblmc_b200.so with the rest of the C ABI (x86-64 baseline: no FMA
// contraction, which the bit-exactness of r2 = x1*x1 + x2*x2 and loc + scale * g relies on).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>
#include "../../include/lmc_b200.h"

namespace {

constexpr int kN = 624, kM = 397;
constexpr uint32_t kMatrixA = 0x9908b0dfu, kUpper = 0x80000000u, kLower = 0x7fffffffu;

struct Mt {
    uint32_t* key;
    int pos;
    uint32_t tw[kN];                                               // the current block, tempered (one vectorisable pass per 624 words)
    void temper_block() {
        for (int i = 0; i < kN; ++i) {
            uint32_t y = key[i];
            y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
            tw[i] = y;
        }
    }
    void regen() {
        int i;
        uint32_t y;
        for (i = 0; i < kN - kM; ++i) { y = (key[i] & kUpper) | (key[i + 1] & kLower); key[i] = key[i + kM] ^ (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA); }
        for (; i < kN - 1; ++i)       { y = (key[i] & kUpper) | (key[i + 1] & kLower); key[i] = key[i + (kM - kN)] ^ (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA); }
        y = (key[kN - 1] & kUpper) | (key[0] & kLower);
        key[kN - 1] = key[kM - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA);
        pos = 0;
        temper_block();
    }
    inline uint32_t next() {
        if (pos == kN) regen();
        return tw[pos++];
    }
    inline double next_double() {                                  // legacy_double: 53 random bits
        const int32_t a = (int32_t)(next() >> 5), b = (int32_t)(next() >> 6);
        return (a * 67108864.0 + b) / 9007199254740992.0;
    }
    inline void pair(double& x1, double& x2) {                     // the rejection loop of legacy_gauss
        double r2;
        do {
            x1 = 2.0 * next_double() - 1.0;
            x2 = 2.0 * next_double() - 1.0;
            r2 = x1 * x1 + x2 * x2;
        } while (r2 >= 1.0 || r2 == 0.0);
    }
};

inline double polar_factor(double x1, double x2) {
    const double r2 = x1 * x1 + x2 * x2;
    return std::sqrt(-2.0 * std::log(r2) / r2);
}

}  // namespace

extern "C" int lmc_host_legacy_normal(uint32_t* mt_key, int32_t* mt_pos, int32_t* has_gauss, double* cached_gauss,
                                      double loc, double scale, int64_t n, double* out, int32_t n_threads) {
    if (!mt_key || !mt_pos || !has_gauss || !cached_gauss || n < 0 || (n > 0 && !out)) return LMC_ERR_INVALID;
    if (*mt_pos < 0 || *mt_pos > kN) return LMC_ERR_INVALID;
    Mt mt;
    mt.key = mt_key; mt.pos = *mt_pos;
    mt.temper_block();
    int64_t i = 0;
    if (n > 0 && *has_gauss) {                                      // the second half of an earlier pair comes first
        out[i++] = loc + scale * *cached_gauss;
        *has_gauss = 0; *cached_gauss = 0.0;
    }
    const int64_t first = i, n_pairs = (n - first) / 2;
    // sequential: accepted pairs, parked in the output itself (returned first: f * x2, then the kept f * x1)
    for (int64_t p = 0; p < n_pairs; ++p) { double x1, x2; mt.pair(x1, x2); out[first + 2 * p] = x2; out[first + 2 * p + 1] = x1; }
    // parallel: the transcendental part
    auto body = [&](int64_t b, int64_t e) {
        for (int64_t p = b; p < e; ++p) {
            double* o = out + first + 2 * p;
            const double x2 = o[0], x1 = o[1], f = polar_factor(x1, x2);
            o[0] = loc + scale * (f * x2);
            o[1] = loc + scale * (f * x1);
        }
    };
    int th = n_threads < 1 ? 1 : (n_threads > 64 ? 64 : n_threads);
    const int64_t min_chunk = 1 << 14;
    if ((n_pairs + min_chunk - 1) / min_chunk < th) th = (int)((n_pairs + min_chunk - 1) / min_chunk);
    if (th <= 1) body(0, n_pairs);
    else {
        std::vector<std::thread> pool;
        pool.reserve(th - 1);
        const int64_t per = (n_pairs + th - 1) / th;
        for (int t = 1; t < th; ++t) { const int64_t b = per * t, e = b + per < n_pairs ? b + per : n_pairs; if (b < e) pool.emplace_back(body, b, e); }
        body(0, per < n_pairs ? per : n_pairs);
        for (auto& t : pool) t.join();
    }
    if (first + 2 * n_pairs < n) {                                  // odd tail: return f * x2, keep f * x1 for the next call
        double x1, x2;
        mt.pair(x1, x2);
        const double f = polar_factor(x1, x2);
        out[n - 1] = loc + scale * (f * x2);
        *cached_gauss = f * x1; *has_gauss = 1;
    }
    *mt_pos = mt.pos;
    return LMC_OK;
}
