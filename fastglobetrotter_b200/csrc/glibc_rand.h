// glibc_rand.h -- host side of the bit-compatible glibc rand() stream.
//
// The reference draws chunk indices with libc rand() (fastGLOBETROTTERCompanion.c:136-137,
// 1147-1148).  glibc's rand() is the TYPE_3 additive-feedback generator of random_r.c:
//     x[n] = x[n-31] + x[n-3]  (mod 2^32),   rand() = x[n] >> 1
// seeded by srandom_r (Lehmer 16807 fill, 310 outputs discarded).  The recurrence is linear
// over Z/2^32, so the state D draws ahead is a fixed linear map of the current 31-word
// history: with the shift operator E (E^31 = E^28 + 1),  E^D = sum_i c_i E^i  (deg < 31) and
//     x[D+j] = sum_i c_i x[i+j].
// This header provides: the history type, seeding, polynomial jump-ahead, the level tables
// the device generator uses to position every thread, and the coupling to the live libc
// state (so the library consumes the very draws the reference would, and leaves rand() where
// the reference would have left it).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace fgt {

constexpr int kLag = 31;      // degree of the generator
constexpr int kTap = 28;      // x[n] = x[n-31] + x[n-31+28]

struct Poly { uint32_t c[kLag]; };

inline Poly poly_one() { Poly p{}; p.c[0] = 1; return p; }
inline Poly poly_E()   { Poly p{}; p.c[1] = 1; return p; }

// (a*b) mod (E^31 - E^28 - 1), coefficients mod 2^32
inline Poly poly_mul(const Poly &a, const Poly &b) {
    uint32_t w[2 * kLag - 1] = {0};
    for (int i = 0; i < kLag; i++) {
        uint32_t ai = a.c[i];
        if (!ai) continue;
        for (int j = 0; j < kLag; j++) w[i + j] += ai * b.c[j];
    }
    for (int d = 2 * kLag - 2; d >= kLag; d--) {   // E^d = E^(d-3) + E^(d-31)
        w[d - (kLag - kTap)] += w[d];
        w[d - kLag] += w[d];
    }
    Poly r;
    std::memcpy(r.c, w, sizeof r.c);
    return r;
}

inline Poly poly_pow_E(uint64_t D) {           // E^D
    Poly result = poly_one(), base = poly_E();
    while (D) {
        if (D & 1) result = poly_mul(result, base);
        D >>= 1;
        if (D) base = poly_mul(base, base);
    }
    return result;
}

struct History { uint32_t x[kLag]; };          // x[0] oldest (x[n-31]) ... x[30] newest (x[n-1])

// history after D more draws
inline History apply_jump(const Poly &p, const History &h) {
    uint32_t y[2 * kLag - 1];
    std::memcpy(y, h.x, sizeof h.x);
    for (int m = 0; m < kLag - 1; m++) y[kLag + m] = y[m] + y[m + kTap];
    History o;
    for (int j = 0; j < kLag; j++) {
        uint32_t s = 0;
        for (int i = 0; i < kLag; i++) s += p.c[i] * y[i + j];
        o.x[j] = s;
    }
    return o;
}

inline History seeded_history(unsigned seed) {   // == state right after srand(seed)
    int32_t r[kLag];
    if (seed == 0) seed = 1;
    r[0] = (int32_t)seed;
    for (int i = 1; i < kLag; i++) {
        long hi = r[i - 1] / 127773, lo = r[i - 1] % 127773;
        long w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        r[i] = (int32_t)w;
    }
    History h;                                   // glibc: fptr=&r[3], rptr=&r[0]
    for (int i = 0; i < kLag; i++) h.x[i] = (uint32_t)r[(3 + i) % kLag];
    for (int n = 0; n < 310; n++) {              // discard 310
        uint32_t v = h.x[0] + h.x[kTap];
        std::memmove(h.x, h.x + 1, (kLag - 1) * sizeof(uint32_t));
        h.x[kLag - 1] = v;
    }
    return h;
}

inline uint32_t step(History &h) {               // one draw (host, slow path)
    uint32_t v = h.x[0] + h.x[kTap];
    std::memmove(h.x, h.x + 1, (kLag - 1) * sizeof(uint32_t));
    h.x[kLag - 1] = v;
    return v >> 1;
}

// ---- coupling to the live libc generator -------------------------------------------------
// initstate()/setstate() hand back a pointer to libc's own state array; word 0 encodes
// (rear_index * 5 + type), words 1..31 are the table.  *fptr (= table[(rear+3)%31]) is x[n-31].
inline bool libc_snapshot(History &h) {
    static char scratch[128];
    char *old = initstate(1u, scratch, sizeof scratch);
    int32_t *o = reinterpret_cast<int32_t *>(old);
    int type = o[0] % 5, rear = o[0] / 5;
    setstate(old);
    if (type != 3) return false;
    const int32_t *tbl = o + 1;
    int f = (rear + 3) % kLag;
    for (int i = 0; i < kLag; i++) h.x[i] = (uint32_t)tbl[(f + i) % kLag];
    return true;
}

inline bool libc_store(const History &h) {
    static char scratch[128];
    char *old = initstate(1u, scratch, sizeof scratch);
    int32_t *o = reinterpret_cast<int32_t *>(old);
    int type = o[0] % 5;
    if (type == 3) {
        for (int i = 0; i < kLag; i++) o[1 + (3 + i) % kLag] = (int32_t)h.x[i];
        o[0] = 5 * 0 + 3;                        // rear = 0  -> fptr = 3
    }
    setstate(old);
    return type == 3;
}

// ---- tables for the device generator -----------------------------------------------------
// A device thread with linear index t (runs of kRun draws each) reaches its start state by
// applying, for every base-256 digit of t, the polynomial  E^(kRun * 256^level * digit).
constexpr int kRun = 1984;                       // draws per generator thread (= 16 * lcm(31,4))
constexpr int kLevels = 4;                       // 2^32 threads * 1984 draws: ample

inline std::vector<uint32_t> build_level_tables() {
    // layout [level][i][digit]: coefficient i of digit d at tab[(level*31 + i)*256 + d]
    std::vector<uint32_t> tab((size_t)kLevels * kLag * 256);
    Poly unit = poly_pow_E(kRun);
    for (int lv = 0; lv < kLevels; lv++) {
        Poly cur = poly_one();
        for (int d = 0; d < 256; d++) {
            for (int i = 0; i < kLag; i++) tab[((size_t)lv * kLag + i) * 256 + d] = cur.c[i];
            cur = poly_mul(cur, unit);
        }
        unit = cur;                              // unit^256
    }
    return tab;
}

// Jump table of the generator that runs inside the pair-evaluation kernel: coefficient i of
// E^(digit * 256^level) at tab[(level*256 + digit)*32 + i] (i < 31; word 31 is padding), raw-draw granularity.
constexpr int kJumpLevelsHost = 6;               // 2^48 raws

inline std::vector<uint32_t> build_jump_table() {
    std::vector<uint32_t> tab((size_t)kJumpLevelsHost * 256 * 32, 0u);
    Poly unit = poly_E();
    for (int lv = 0; lv < kJumpLevelsHost; lv++) {
        Poly cur = poly_one();
        for (int d = 0; d < 256; d++) {
            for (int i = 0; i < kLag; i++) tab[((size_t)lv * 256 + d) * 32 + i] = cur.c[i];
            cur = poly_mul(cur, unit);
        }
        unit = cur;                              // unit^256
    }
    return tab;
}

}  // namespace fgt
