// speck_b200/host/rmat_gen.cpp -- fast R-MAT edge generator for the synthetic workloads (bench / tests).
// Bit-identical to the numpy definition in speck_b200/matrices.py::rmat: the caller
// passes the state of numpy's PCG64 stream (default_rng(seed).bit_generator.state) and this file replays it --
// per level i: m doubles for the row bit, then m doubles for the column bit -- so the matrix is the one the numpy
// loop produces, only ~20x faster (R-MAT scale 24: 230 s -> ~15 s).  Also sorts the keys and drops duplicates.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {
typedef unsigned __int128 u128;

struct Pcg64 {
    u128 state, inc;
    inline uint64_t next()
    {
        const u128 mult = ((u128)0x2360ED051FC65DA4ull << 64) | 0x4385DF649FCCF645ull;
        state = state * mult + inc;
        const uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state;
        const uint64_t x = hi ^ lo;
        const unsigned rot = (unsigned)(hi >> 58);
        return (x >> rot) | (x << ((64 - rot) & 63));
    }
    inline double next_double() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

// LSD radix sort of 64-bit keys below 2^bits, 16 bits per pass
void radix_sort(std::vector<uint64_t> &a, int bits)
{
    std::vector<uint64_t> b(a.size());
    for (int shift = 0; shift < bits; shift += 16) {
        std::vector<size_t> cnt(65537, 0);
        for (uint64_t k : a) ++cnt[((k >> shift) & 0xffff) + 1];
        for (int i = 0; i < 65536; ++i) cnt[i + 1] += cnt[i];
        for (uint64_t k : a) b[cnt[(k >> shift) & 0xffff]++] = k;
        a.swap(b);
    }
}
}  // namespace

extern "C" {

// keys_out: caller-allocated m entries; returns the number of distinct sorted keys (row * 2^scale + col) written
uint64_t speck_host_rmat_keys(uint64_t state_hi, uint64_t state_lo, uint64_t inc_hi, uint64_t inc_lo, int scale,
                              uint64_t m, double ab, double c_norm, double a_norm, uint64_t *keys_out)
{
    Pcg64 g{((u128)state_hi << 64) | state_lo, ((u128)inc_hi << 64) | inc_lo};
    std::vector<uint64_t> key(m, 0);
    std::vector<uint8_t> ii(m);
    for (int lvl = 0; lvl < scale; ++lvl) {
        for (uint64_t e = 0; e < m; ++e) ii[e] = g.next_double() > ab;
        const uint64_t rbit = 1ull << (scale + lvl), cbit = 1ull << lvl;
        for (uint64_t e = 0; e < m; ++e) {
            const bool jj = g.next_double() > (ii[e] ? c_norm : a_norm);
            key[e] |= (ii[e] ? rbit : 0ull) | (jj ? cbit : 0ull);
        }
    }
    radix_sort(key, 2 * scale);
    uint64_t n = 0;
    for (uint64_t e = 0; e < m; ++e)
        if (e == 0 || key[e] != key[e - 1]) keys_out[n++] = key[e];
    return n;
}

}  // extern "C"
