// Scene GENERATION randomness of three examples (examples/big-scene.rs:27,
// graphics-castle.rs:411, graphics-temple.rs:444): rand 0.7.0 `StdRng` =
// ChaCha20 (rand_chacha 0.2.0), `seed_from_u64` (rand_core 0.5 PCG32 key
// expansion), `gen::<f64>()`, `gen_range`, `SliceRandom::{choose, shuffle}`.
// The crates' sources are not available offline; this is a restatement of
// their published algorithms (SURVEY §2a "parity unpinned").  The visual pin
// is tests/test_reference_renders.py: big-scene rendered through this
// generator is compared with the reference's own render/09a_kdtree.png.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace portrayer {

class StdRng {
  public:
    static StdRng seed_from_u64(uint64_t state) {
        // rand_core 0.5 SeedableRng::seed_from_u64: PCG32 stream fills the 32-byte seed
        const uint64_t MUL = 6364136223846793005ull, INC = 11634580027462260723ull;
        uint8_t seed[32];
        for (int i = 0; i < 8; ++i) {
            state = state * MUL + INC;
            uint32_t xorshifted = static_cast<uint32_t>(((state >> 18) ^ state) >> 27);
            uint32_t rot = static_cast<uint32_t>(state >> 59);
            uint32_t x = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
            std::memcpy(seed + 4 * i, &x, 4);  // to_le on a little-endian host
        }
        return StdRng(seed);
    }

    uint32_t next_u32() {
        if (index_ >= kBuf) generate_and_set(0);
        return results_[index_++];
    }

    // rand_core BlockRng::next_u64
    uint64_t next_u64() {
        if (index_ < kBuf - 1) {
            uint64_t lo = results_[index_], hi = results_[index_ + 1];
            index_ += 2;
            return (hi << 32) | lo;
        } else if (index_ >= kBuf) {
            generate_and_set(2);
            return (static_cast<uint64_t>(results_[1]) << 32) | results_[0];
        } else {
            uint64_t x = results_[kBuf - 1];
            generate_and_set(1);
            uint64_t y = results_[0];
            return (y << 32) | x;
        }
    }

    // Standard f64: 53 random bits * 2^-53, [0,1)
    double gen_f64() { return static_cast<double>(next_u64() >> 11) * (1.0 / 9007199254740992.0); }

    // gen_range(low, high) for usize: UniformInt::sample_single (widening multiply, conservative zone)
    uint64_t gen_range(uint64_t low, uint64_t high) {
        if (!(low < high)) throw std::runtime_error("UniformSampler::sample_single: low >= high");
        const uint64_t range = high - low;
        const int lz = __builtin_clzll(range);
        const uint64_t zone = (range << lz) - 1;
        for (;;) {
            const uint64_t v = next_u64();
            const unsigned __int128 wide = static_cast<unsigned __int128>(v) * range;
            const uint64_t hi = static_cast<uint64_t>(wide >> 64), lo = static_cast<uint64_t>(wide);
            if (lo <= zone) return low + hi;
        }
    }

    // gen_range(low, high) for u32: same scheme on 32-bit words (one next_u32 per try)
    uint32_t gen_range_u32(uint32_t low, uint32_t high) {
        if (!(low < high)) throw std::runtime_error("UniformSampler::sample_single: low >= high");
        const uint32_t range = high - low;
        const uint32_t zone = (range << __builtin_clz(range)) - 1;
        for (;;) {
            const uint32_t v = next_u32();
            const uint64_t wide = static_cast<uint64_t>(v) * range;
            const uint32_t hi = static_cast<uint32_t>(wide >> 32), lo = static_cast<uint32_t>(wide);
            if (lo <= zone) return low + hi;
        }
    }

    // rand 0.7 seq::gen_index: sample a u32 when the bound fits (value-stable across 32/64-bit hosts)
    size_t gen_index(size_t ubound) {
        if (ubound <= 0xFFFFFFFFull) return gen_range_u32(0, static_cast<uint32_t>(ubound));
        return gen_range(0, ubound);
    }

    // SliceRandom::choose
    template <class T>
    const T& choose(const std::vector<T>& v) {
        if (v.empty()) throw std::runtime_error("choose on empty slice");
        return v[gen_index(v.size())];
    }

    // SliceRandom::shuffle: for i in (1..len).rev() { swap(i, gen_index(i + 1)) }
    template <class T>
    void shuffle(std::vector<T>& v) {
        for (size_t i = v.size(); i-- > 1;) std::swap(v[i], v[gen_index(i + 1)]);
    }

    // Rng::gen_bool is not used by the examples; gen::<bool>() = (next_u32() as i32) < 0
    bool gen_bool_standard() { return static_cast<int32_t>(next_u32()) < 0; }

  private:
    static constexpr size_t kBuf = 64;  // rand_chacha 0.2 refills four 16-word blocks at a time
    uint32_t key_[8];
    uint64_t counter_ = 0;
    uint32_t results_[kBuf];
    size_t index_ = kBuf;

    explicit StdRng(const uint8_t seed[32]) { std::memcpy(key_, seed, 32); }

    static uint32_t rotl(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }
    static void quarter(uint32_t* x, int a, int b, int c, int d) {
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 16);
        x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 12);
        x[a] += x[b]; x[d] = rotl(x[d] ^ x[a], 8);
        x[c] += x[d]; x[b] = rotl(x[b] ^ x[c], 7);
    }
    void block(uint64_t counter, uint32_t* out) const {
        uint32_t st[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
        for (int i = 0; i < 8; ++i) st[4 + i] = key_[i];
        st[12] = static_cast<uint32_t>(counter);
        st[13] = static_cast<uint32_t>(counter >> 32);
        st[14] = 0;  // stream id 0
        st[15] = 0;
        uint32_t x[16];
        std::memcpy(x, st, sizeof x);
        for (int round = 0; round < 10; ++round) {  // 20 rounds = 10 double rounds
            quarter(x, 0, 4, 8, 12); quarter(x, 1, 5, 9, 13); quarter(x, 2, 6, 10, 14); quarter(x, 3, 7, 11, 15);
            quarter(x, 0, 5, 10, 15); quarter(x, 1, 6, 11, 12); quarter(x, 2, 7, 8, 13); quarter(x, 3, 4, 9, 14);
        }
        for (int i = 0; i < 16; ++i) out[i] = x[i] + st[i];
    }
    void generate_and_set(size_t index) {
        for (int b = 0; b < 4; ++b) block(counter_ + b, results_ + 16 * b);
        counter_ += 4;
        index_ = index;
    }
};

}  // namespace portrayer
