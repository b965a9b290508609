// TEST INFRASTRUCTURE ONLY (oracle/_ref build): C entry points over the UNMODIFIED reference prover
// (src/prover.cpp compiled with -Dprover=ref_prover) for circuits handed over in memory. Used as the
// CPU baseline of bench.py (`kind: "reference"`) and to cross-check the oracle on circuits that have
// no .pws form. The driver below calls the prover in the order of verifier::verify
// (src/verifier.cpp:134-337) and draws the challenges with the reference's own F::random(); it does
// not re-check the rounds (the oracle's verifier does that in the tests) and skips the polynomial
// commitment.
#include <chrono>
#include <cmath>
#include <cstdint>
#include <vector>

#include "circuit.h"
#include "config_pc.hpp"
#include "polynomial.h"
#include "utils.hpp"

#define prover ref_prover
#include REF_PROVER_H
#undef prover

struct ofe { uint64_t re, im; };
static inline ofe to_ofe(const F &x) { return ofe{x.real, x.img}; }

// cst / is_assert (per gate, layer 0 included; may be null): gate constants of Addc / Mulc gates (any F, also complex) and
// assert flags -- gate types and fields the .pws parser never produces but prover.cpp handles (prover.cpp:60-83,211,229-272).
extern "C" int ref_gkr_prove2(int n_layers, const uint64_t *layer_size, const uint8_t *ty, const int32_t *l,
                              const uint32_t *u, const uint32_t *v, const uint64_t *inputs, const ofe *cst,
                              const uint8_t *is_assert, ofe *tr, double *prove_seconds, double *eval_seconds);
extern "C" int ref_gkr_prove(int n_layers, const uint64_t *layer_size, const uint8_t *ty, const int32_t *l,
                             const uint32_t *u, const uint32_t *v, const uint64_t *inputs, unsigned seed,
                             ofe *tr, double *prove_seconds, double *eval_seconds) {
    (void)seed;  // F::init() seeds 3396 (fieldElement.cpp:108)
    return ref_gkr_prove2(n_layers, layer_size, ty, l, u, v, inputs, nullptr, nullptr, tr, prove_seconds, eval_seconds);
}
extern "C" int ref_gkr_prove2(int n_layers, const uint64_t *layer_size, const uint8_t *ty, const int32_t *l,
                              const uint32_t *u, const uint32_t *v, const uint64_t *inputs, const ofe *cst,
                              const uint8_t *is_assert, ofe *tr, double *prove_seconds, double *eval_seconds) {
    layeredCircuit c;
    c.size = n_layers;
    c.circuit.resize(n_layers);
    size_t off = 0;
    for (int i = 0; i < n_layers; ++i) {
        layer &L = c.circuit[i];
        L.size = layer_size[i];
        L.gates.resize(L.size);
        for (u64 g = 0; g < L.size; ++g) {
            if (i == 0) L.gates[g] = gate(gateType::Input, -1, inputs[g], 0, F_ZERO, false);
            else {
                F cc = F_ZERO;
                if (cst) { cc.real = cst[off + g].re; cc.img = cst[off + g].im; }
                L.gates[g] = gate((gateType)ty[off + g], l[off + g], u[off + g], v[off + g], cc, is_assert && is_assert[off + g]);
            }
        }
        L.bitLength = (int)log2(L.size);  // main.cpp:133-136
        if ((1ULL << L.bitLength) < L.size) ++L.bitLength;
        off += L.size;
    }
    F::init();
    c.subsetInit();
    auto t0 = std::chrono::steady_clock::now();
    ref_prover p(c);  // runs evaluate()
    double ev = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    p.init();

    int max_bl = 0;
    for (auto &L : c.circuit) max_bl = std::max(max_bl, L.bitLength);
    std::vector<F> r_u(max_bl), r_liu(max_bl), sig(n_layers);
    std::vector<std::vector<F>> r_v(n_layers), claims_v(n_layers);
    size_t ti = 0;
    auto put = [&](const F &x) { tr[ti++] = to_ofe(x); };
    auto putq = [&](const quadratic_poly &q) { put(q.a); put(q.b); put(q.c); };

    const int out_bl = c.circuit[n_layers - 1].bitLength;
    for (int i = 0; i < out_bl; ++i) r_liu[i] = F::random();
    put(p.Vres(r_liu.begin(), out_bl));
    p.sumcheckInitAll(r_liu.begin());
    for (int i = n_layers - 1; i; --i) {
        const int pb = c.circuit[i - 1].bitLength, m = c.circuit[i].maxDadBitLength;
        p.sumcheckInit();
        for (auto &x : r_u) x = F::random();
        F assert_random = F::random();
        p.sumcheckInitPhase1(assert_random);
        F prev = F_ZERO, claim_u;
        for (int j = 0; j < pb; ++j) { putq(p.sumcheckUpdatePhase1(prev)); prev = r_u[j]; }
        p.sumcheckFinalize1(prev, claim_u);
        put(claim_u);
        if (~m) {
            r_v[i].resize(m);
            claims_v[i].assign(i, F_ZERO);
            for (auto &x : r_v[i]) x = F::random();
            p.sumcheckInitPhase2();
            prev = F_ZERO;
            for (int j = 0; j < m; ++j) { putq(p.sumcheckUpdatePhase2(prev)); prev = r_v[i][j]; }
            p.sumcheckFinalize2(prev, claims_v[i].begin());
            for (int s = 0; s < i; ++s) put(claims_v[i][s]);
        }
        for (auto &x : sig) x = F::random();
        for (auto &x : r_liu) x = F::random();
        p.sumcheckInitLiu(sig.begin());
        prev = F_ZERO;
        for (int j = 0; j < pb; ++j) { putq(p.sumcheckLiuUpdate(prev)); prev = r_liu[j]; }
        F vr;
        p.sumcheckLiuFinalize(prev, vr);
        put(vr);
    }
    // verifier.cpp:367-369 + prover.cpp:544: <circuitValue[0], eq(r_liu,.)>
    {
        const int b0 = c.circuit[0].bitLength;
        std::vector<F> eq(1ULL << b0), in0(1ULL << b0, F_ZERO);
        initBetaTable(eq, b0, r_liu.begin(), F_ONE);
        for (u64 g = 0; g < c.circuit[0].size; ++g) in0[g] = F((long long)c.circuit[0].gates[g].u);
        put(p.inner_prod(in0, eq, c.circuit[0].size));
    }
    if (prove_seconds) *prove_seconds = p.proveTime();
    if (eval_seconds) *eval_seconds = ev;
    return (int)ti;
}

// The polynomial commitment is never exercised through this library; the prebuilt libXKCP.a holds
// non-PIC objects that cannot go into a shared object, so its one entry point used by the
// reference (my_hhash.h:29) is stubbed to trap.
extern "C" int SHA3_256(unsigned char *, const unsigned char *, size_t) { __builtin_trap(); }
