// Drop-in replacement for the reference's src/prover.h (/root/reference/src/prover.h:12-67):
// the same class name, constructor and public methods, so the reference's verifier.cpp / main.cpp
// compile and link against it unchanged. Every method forwards to the C ABI of the B200 engine
// (include/virgo_b200.h); no prover arithmetic happens on the host and there is no CPU fallback --
// a CUDA failure prints the error and exits, like the reference's own failure path
// (prover.cpp:18-21: fprintf(stderr) + exit(EXIT_FAILURE)).
//
// Build: place this file where the reference's prover.h is found first, i.e. the
// includes below resolve to the REFERENCE's own circuit.h / config_pc.hpp / polynomial.h.
#pragma once

#include <vector>

#include "circuit.h"
#include "config_pc.hpp"
#include "polynomial.h"

struct vp_ctx;
struct vp_circuit;

class prover {
public:
    // prover.cpp:14-25 -- builds the device context (vp_create / vp_create_sharded) and, like the reference's
    // constructor, evaluates the circuit
    explicit prover(const layeredCircuit &cir);
    ~prover();
    prover(const prover &) = delete;
    prover &operator=(const prover &) = delete;

    // ---- circuit values and bookkeeping
    void evaluate();                                                    // prover.cpp:27-97    -> vp_evaluate
    F Vres(const vector<F>::const_iterator &r_0, int r_0_size);         // prover.cpp:99-129   -> vp_vres
    void init();                                                        // prover.cpp:131-160  (state lives in the context)
    void sumcheckInitAll(const vector<F>::const_iterator &r_last);      // prover.cpp:162-175  -> vp_sumcheck_init_all
    void sumcheckInit();                                                // prover.cpp:177-187  -> vp_sumcheck_init

    // ---- the three sumchecks of a layer: table set-up, one round per call, final claims
    void sumcheckInitPhase1(const F &assert_random);                    // prover.cpp:189-280  -> vp_init_phase1
    quadratic_poly sumcheckUpdatePhase1(const F &previousRandom);       // prover.cpp:422-425  -> vp_round(1, ..)
    void sumcheckFinalize1(const F &previousRandom, F &claim);          // prover.cpp:494-502  -> vp_finalize1

    void sumcheckInitPhase2();                                          // prover.cpp:282-367  -> vp_init_phase2
    quadratic_poly sumcheckUpdatePhase2(const F &previousRandom);       // prover.cpp:427-430  -> vp_round(2, ..)
    void sumcheckFinalize2(const F &previousRandom, vector<F>::iterator claims);   // prover.cpp:504-516 -> vp_finalize2

    void sumcheckInitLiu(vector<F>::const_iterator s);                  // prover.cpp:369-420  -> vp_init_liu
    quadratic_poly sumcheckLiuUpdate(const F &previousRandom);          // prover.cpp:432-434  -> vp_round(3, ..)
    void sumcheckLiuFinalize(const F &previousRandom, F &claim);        // prover.cpp:518-522  -> vp_finalize_liu

    // ---- statistics the reference's main.cpp prints
    double proveTime() const;                                           // prover.cpp:549-551  -> vp_prove_seconds
    double proofSize() const;                                           // prover.cpp:553-555  -> vp_proof_size_bytes (kB)

#ifdef USE_VIRGO
    // ---- polynomial commitment of the input layer (the reference's verifier calls these, verifier.cpp:352-383)
    virgo::poly_commit::poly_commit_prover poly_prover;                 // the reference's CPU opening phase keeps using it
    virgo::__hhash_digest commit_private();                             // prover.cpp:524-530  -> vp_commit_private
    F inner_prod(const vector<F> &a, const vector<F> &b, u64 l);        // prover.cpp:532-540  -> vp_inner_prod
    virgo::__hhash_digest commit_public(vector<F> &pub, F &inner_product_sum, std::vector<F> &mask,
                                        vector<F> &all_sum);            // prover.cpp:542-547  -> vp_commit_public
#endif

private:
    quadratic_poly round(int phase, const F &previousRandom);

    const layeredCircuit &C;
    vp_circuit *circ;
    vp_ctx *ctx;
    int sumcheckLayerId;
    int world;                    // GPUs this prover is sharded over (VP_WORLD copies of the program, one per GPU)
    bool gpu_commit;              // commit_private ran on the device: commit_public does too
    std::vector<F> input_values;  // host copy of circuitValue[0], zero-padded to 2^bitLength (for the PC)
};
