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
    explicit prover(const layeredCircuit &cir);
    ~prover();
    prover(const prover &) = delete;
    prover &operator=(const prover &) = delete;

    void evaluate();
    void init();
    void sumcheckInitAll(const vector<F>::const_iterator &r_last);
    void sumcheckInit();
    void sumcheckInitPhase1(const F &assert_random);
    void sumcheckInitPhase2();
    void sumcheckInitLiu(vector<F>::const_iterator s);

    quadratic_poly sumcheckUpdatePhase1(const F &previousRandom);
    quadratic_poly sumcheckUpdatePhase2(const F &previousRandom);
    quadratic_poly sumcheckLiuUpdate(const F &previousRandom);

    void sumcheckFinalize1(const F &previousRandom, F &claim);
    void sumcheckFinalize2(const F &previousRandom, vector<F>::iterator claims);
    void sumcheckLiuFinalize(const F &previousRandom, F &claim);

    F Vres(const vector<F>::const_iterator &r_0, int r_0_size);

    double proveTime() const;
    double proofSize() const;

#ifdef USE_VIRGO
    virgo::poly_commit::poly_commit_prover poly_prover;
    virgo::__hhash_digest commit_private();
    F inner_prod(const vector<F> &a, const vector<F> &b, u64 l);
    virgo::__hhash_digest commit_public(vector<F> &pub, F &inner_product_sum, std::vector<F> &mask, vector<F> &all_sum);
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
