// sketch.cpp — CPU ORACLE (test infrastructure): recode tables + SplitMix64 + Bloom + CMS.
// Restates /root/reference/src/recode.rs and /root/reference/src/proba_filter.rs.
#include "kamino_oracle.hpp"
#include <cstdlib>
#include <cstring>

namespace kmo {

// ---------------------------------------------------------------- recode.rs
static uint8_t lower_to_upper(uint8_t b) { return (b >= 'a' && b <= 'z') ? uint8_t(b - 32) : b; }

// src/recode.rs:40-51 — both cases of each letter map to the class.
static uint8_t dayhoff6_code(uint8_t b) {
    switch (lower_to_upper(b)) {
        case 'C': return 0;
        case 'A': case 'G': case 'P': case 'S': case 'T': return 1;
        case 'D': case 'E': case 'N': case 'Q': return 2;
        case 'H': case 'K': case 'R': return 3;
        case 'I': case 'L': case 'M': case 'V': return 4;
        case 'F': case 'W': case 'Y': return 5;
        default: return 255;
    }
}
// src/recode.rs:54-65
static uint8_t sr6_code(uint8_t b) {
    switch (lower_to_upper(b)) {
        case 'A': case 'P': case 'S': case 'T': return 0;
        case 'D': case 'E': case 'N': case 'G': return 1;
        case 'Q': case 'K': case 'R': return 2;
        case 'M': case 'I': case 'V': case 'L': return 3;
        case 'W': case 'C': return 4;
        case 'F': case 'Y': case 'H': return 5;
        default: return 255;
    }
}
// src/recode.rs:68-80
static uint8_t kgb6_code(uint8_t b) {
    switch (lower_to_upper(b)) {
        case 'A': case 'G': case 'P': case 'S': return 0;
        case 'D': case 'E': case 'N': case 'Q': case 'H': case 'K': case 'R': case 'T': return 1;
        case 'M': case 'I': case 'L': return 2;
        case 'W': return 3;
        case 'F': case 'Y': return 4;
        case 'C': case 'V': return 5;
        default: return 255;
    }
}
// src/recode.rs:31-37.  NB lower_to_upper only folds a-z, and every arm of the reference
// lists exactly the upper- and lower-case letter, so folding first is equivalent.
uint8_t recode_byte(uint8_t b, RecodeScheme s) {
    if (b >= 128) return 255;
    switch (s) {
        case Dayhoff6: return dayhoff6_code(b);
        case SR6: return sr6_code(b);
        default: return kgb6_code(b);
    }
}
const char* scheme_name(RecodeScheme s) {   // src/recode.rs:18-28
    return s == Dayhoff6 ? "Dayhoff6" : (s == SR6 ? "SR6" : "KGB6");
}

// ---------------------------------------------------------- proba_filter.rs
const uint64_t SEEDS[6] = {  // src/proba_filter.rs:21-28
    0x9e3779b97f4a7c15ull, 0xbf58476d1ce4e5b9ull, 0x94d049bb133111ebull,
    0x3c79ac492ba7b653ull, 0x1c69b3f74ac4ae35ull, 0x6a09e667f3bcc909ull,
};

uint64_t mix64(uint64_t x) {  // src/proba_filter.rs:30-37
    x ^= x >> 30;
    x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27;
    x *= 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

static size_t next_pow2(size_t v) {
    size_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

BloomFilter::BloomFilter(size_t nbits) {  // src/proba_filter.rs:48-56
    size_t size = next_pow2(nbits < 64 ? 64 : nbits);
    size_t words = (size + 63) / 64;
    bits.assign(words, 0);
    mask = uint64_t(size - 1);
}
void BloomFilter::clear() { std::fill(bits.begin(), bits.end(), 0); }

bool BloomFilter::test_and_set(uint64_t key) {  // src/proba_filter.rs:59-74
    uint64_t h1 = mix64(key ^ SEEDS[0]);
    uint64_t h2 = mix64(key ^ SEEDS[1]) | 1;
    bool seen = true;
    for (uint64_t i = 0; i < BLOOM_HASHES; i++) {
        size_t idx = size_t((h1 + i * h2) & mask);
        size_t word = idx >> 6;
        uint64_t bit = uint64_t(1) << (idx & 63);
        bool had = (bits[word] & bit) != 0;
        bits[word] |= bit;
        if (!had) seen = false;
    }
    return seen;
}

static std::vector<uint64_t> cms_seeds(size_t d) {  // src/proba_filter.rs:91-95,131-135
    std::vector<uint64_t> s;
    for (size_t i = 0; i < d; i++) s.push_back(SEEDS[2 + (i % (6 - 2))]);
    return s;
}

CountMinSketch::CountMinSketch(size_t w_, size_t d_) {  // src/proba_filter.rs:88-101
    size_t w = next_pow2(w_ < 1 ? 1 : w_);
    size_t d = d_ < 1 ? 1 : d_;
    width = w;
    seeds = cms_seeds(d);
    table.assign(w * d, 0);
}

uint32_t CountMinSketch::estimate(uint64_t key) const {  // src/proba_filter.rs:103-112
    uint16_t mn = 0xFFFF;
    for (size_t row = 0; row < seeds.size(); row++) {
        size_t idx = size_t(mix64(key ^ seeds[row]) & (uint64_t(width) - 1));
        uint16_t v = table[row * width + idx];
        if (v < mn) mn = v;
    }
    return mn;
}

AtomicCountMinSketch::AtomicCountMinSketch(size_t w_, size_t d_) {  // src/proba_filter.rs:129-143
    size_t w = next_pow2(w_ < 1 ? 1 : w_);
    size_t d = d_ < 1 ? 1 : d_;
    width = w;
    seeds = cms_seeds(d);
    cells = w * d;
    table = static_cast<uint16_t*>(std::calloc(cells, sizeof(uint16_t)));
    if (!table) throw std::runtime_error("oracle: cannot allocate count-min sketch");
}
AtomicCountMinSketch::~AtomicCountMinSketch() { std::free(table); }

void AtomicCountMinSketch::increment(uint64_t key) {  // src/proba_filter.rs:146-168
    for (size_t row = 0; row < seeds.size(); row++) {
        size_t idx = size_t(mix64(key ^ seeds[row]) & (uint64_t(width) - 1));
        uint16_t* cell = &table[row * width + idx];
        uint16_t current = __atomic_load_n(cell, __ATOMIC_RELAXED);
        for (;;) {
            if (current == 0xFFFF) break;
            uint16_t next = uint16_t(current + 1);
            if (__atomic_compare_exchange_n(cell, &current, next, true, __ATOMIC_RELAXED,
                                            __ATOMIC_RELAXED))
                break;
        }
    }
}

CountMinSketch AtomicCountMinSketch::snapshot() const {  // src/proba_filter.rs:171-177
    CountMinSketch cms(width, seeds.size());
    for (size_t i = 0; i < cells; i++) cms.table[i] = __atomic_load_n(&table[i], __ATOMIC_RELAXED);
    return cms;
}

}  // namespace kmo
