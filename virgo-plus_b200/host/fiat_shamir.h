// Fiat-Shamir challenge source: the reference ships `transcriptCache`
// (/root/reference/lib/virgo/src/transcriptCache.hpp:14-50) -- a byte pool that is hashed with SHA3-256 whenever a
// challenge is requested -- but never calls it (its verifier draws glibc random() values that do not depend on the
// prover's messages). This is that class restated, plus the one thing the dead code leaves open: WHEN
// messages are stored and challenges drawn (fs_order below).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <vector>

#include "../csrc/field.cuh"

namespace vp {

void sha3_256(const unsigned char* msg, size_t len, unsigned char out[32]);   // FIPS 202 (XKCP's SHA3_256 in the reference)

class FsCache {   // transcriptCache.hpp:14-50
public:
    void store(const void* in, size_t n) { const unsigned char* p = (const unsigned char*)in; pool.insert(pool.end(), p, p + n); }
    void store(const F& x) { store(&x, sizeof x); }   // 16 bytes {u64 real, u64 img}, like store(const T&) on a fieldElement
    F random() {   // :40-46: hash the pool, the digest becomes the pool, first two words mod p
        unsigned char out[32];
        sha3_256(pool.data(), pool.size(), out);
        pool.assign(out, out + 32);
        uint64_t re, im;
        __builtin_memcpy(&re, out, 8);
        __builtin_memcpy(&im, out + 8, 8);
        return F{re % P, im % P};
    }
private:
    std::vector<unsigned char> pool;
};

// Order of stores and draws in Fiat-Shamir mode (a sound variant of the interactive order of verifier.cpp:134-337: a round's
// challenge is drawn AFTER the round's polynomial was stored; the reference's interactive verifier draws a phase's
// challenges before the phase because they are independent of the messages):
//   store(seed[32]);  r_out[0..bl(out)) = draws;  store(Vres);
//   per layer i = n-1..1:  assert_random = draw;
//      phase 1: per round j: store(a, b, c); r_u[j] = draw;                     then store(claim_u);
//      phase 2 (if any): per round j: store(a, b, c); r_v[j] = draw;            then store(claims_v[0..i));
//      sig[0..n) = draws;  Liu: per round j: store(a, b, c); r_liu[j] = draw;   then store(claim_liu);
//   store(input MLE).
// Challenge slots of the usual layout that are never used (r_u[j], r_liu[j] for j >= the phase's rounds) stay zero.

}  // namespace vp
