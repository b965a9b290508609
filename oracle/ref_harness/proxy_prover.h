// TEST INFRASTRUCTURE ONLY (oracle/_ref build). A logging proxy with the reference's `prover`
// interface (src/prover.h:12-42): the UNMODIFIED reference verifier.cpp is compiled against this
// header (through a symlink farm, see oracle/Makefile) and every call is forwarded to the
// UNMODIFIED reference prover, compiled from src/prover.cpp with -Dprover=ref_prover. Every
// prover->verifier message and every challenge the prover receives is appended to a log, in the
// tag order of SURVEY.md 9.5. Nothing here is shipped or used by the product.
#pragma once
#include <vector>

#include "circuit.h"
#include "config_pc.hpp"
#include "polynomial.h"

#define prover ref_prover
#include REF_PROVER_H
#undef prover

void ref_log(const char *tag, const F &x);
// REF_TAMPER=<k> (ref_dump): the k-th prover->verifier message (transcript order: VRES, PA, PB, PC, CLAIM_*, ..., INPUT_MLE;
// challenges are not counted) reaches the UNMODIFIED verifier with 1 added to it, and is logged as the verifier saw it.
F ref_tamper(const F &x);

class prover {
public:
    explicit prover(const layeredCircuit &cir) : ref(cir), poly_prover(ref.poly_prover) {}
    void evaluate() { ref.evaluate(); }
    void init() { ref.init(); }
    void sumcheckInitAll(const vector<F>::const_iterator &r_last) { ref.sumcheckInitAll(r_last); }
    void sumcheckInit() { ref.sumcheckInit(); --cur_layer; }
    void sumcheckInitPhase1(const F &assert_random) { ref.sumcheckInitPhase1(assert_random); }
    void sumcheckInitPhase2() { ref.sumcheckInitPhase2(); }
    void sumcheckInitLiu(vector<F>::const_iterator s) { ref.sumcheckInitLiu(s); }

    quadratic_poly log_poly(const F &prev, quadratic_poly p) {
        ref_log("CH", prev);
        p.a = ref_tamper(p.a);
        p.b = ref_tamper(p.b);
        p.c = ref_tamper(p.c);
        ref_log("PA", p.a);
        ref_log("PB", p.b);
        ref_log("PC", p.c);
        return p;
    }
    quadratic_poly sumcheckUpdatePhase1(const F &prev) { return log_poly(prev, ref.sumcheckUpdatePhase1(prev)); }
    quadratic_poly sumcheckUpdatePhase2(const F &prev) { return log_poly(prev, ref.sumcheckUpdatePhase2(prev)); }
    quadratic_poly sumcheckLiuUpdate(const F &prev) { return log_poly(prev, ref.sumcheckLiuUpdate(prev)); }

    void sumcheckFinalize1(const F &prev, F &claim) {
        ref.sumcheckFinalize1(prev, claim);
        claim = ref_tamper(claim);
        ref_log("CH", prev);
        ref_log("CLAIM_U", claim);
    }
    void sumcheckFinalize2(const F &prev, vector<F>::iterator claims) {
        ref.sumcheckFinalize2(prev, claims);
        ++n_fin2;
        // the verifier hands final_claims_v[layer].begin(), which has exactly `layer` entries; the
        // layer id counts down from size-1 on every sumcheckInit (prover.cpp:166,179)
        for (int i = 0; i < cur_layer; ++i) { claims[i] = ref_tamper(claims[i]); ref_log("CLAIM_V", claims[i]); }
    }
    void sumcheckLiuFinalize(const F &prev, F &claim) {
        ref.sumcheckLiuFinalize(prev, claim);
        claim = ref_tamper(claim);
        ref_log("CH", prev);
        ref_log("CLAIM_LIU", claim);
    }
    F Vres(const vector<F>::const_iterator &r_0, int r_0_size) {
        F x = ref_tamper(ref.Vres(r_0, r_0_size));
        ref_log("VRES", x);
        return x;
    }
    double proveTime() const { return ref.proveTime(); }
    double proofSize() const { return ref.proofSize(); }

    virgo::__hhash_digest commit_private() { return ref.commit_private(); }
    F inner_prod(const vector<F> &a, const vector<F> &b, u64 l) { return ref.inner_prod(a, b, l); }
    virgo::__hhash_digest commit_public(vector<F> &pub, F &sum, std::vector<F> &mask, vector<F> &all_sum) {
        auto d = ref.commit_public(pub, sum, mask, all_sum);
        sum = ref_tamper(sum);
        ref_log("INPUT_MLE", sum);
        return d;
    }

    ref_prover ref;
    virgo::poly_commit::poly_commit_prover &poly_prover;
    int cur_layer = 0;  // harness sets it to C.size; sumcheckInit counts it down like sumcheckLayerId
    int n_fin2 = 0;
};
