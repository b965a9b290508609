/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the Virgo++ GKR prover/verifier path.
 *
 * Plain-C restatement of the reference algorithm (file:line citations are on each function in
 * gkr_oracle.c). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * this library; the product (virgo-plus_b200/) never links or calls it.
 *
 * Parity status: PINNED. The oracle is checked (tests/test_oracle.py) against transcripts produced
 * by the unmodified reference prover + verifier (oracle/_ref, built from /root/reference by
 * oracle/Makefile) -- the committed fixtures under tests/golden/ -- and against the known-answer
 * values of SURVEY.md 9.5.
 */
#ifndef GKR_ORACLE_H
#define GKR_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t re, im; } ofe; /* == virgo::fieldElement {real, img} */

/* Flat layered circuit (one instance, already replicated if needed). */
typedef struct {
    int32_t n_layers;
    const uint64_t* layer_size; /* [n_layers] */
    const uint64_t* gate_off;   /* [n_layers+1] offsets into the gate arrays */
    const uint8_t* ty;          /* gateType values of inputCircuit.hpp:14-16 */
    const int32_t* l;
    const uint32_t* u;
    const uint32_t* v;
    const uint32_t* lv;
    const ofe* c;               /* may be NULL (no Addc/Mulc) */
    const uint8_t* is_assert;   /* may be NULL */
    const uint64_t* dad_size;   /* [n_layers*n_layers], index i*n_layers + l */
    const uint64_t* dad_off;    /* [n_layers*n_layers + 1] offsets into dad_id, same index */
    const uint32_t* dad_id;
    const uint64_t* inputs;     /* [layer_size[0]] */
} ogkr_circuit;

/* field (exposed for the field tests) */
ofe ofe_add(ofe a, ofe b);
ofe ofe_sub(ofe a, ofe b);
ofe ofe_mul(ofe a, ofe b);

/* glibc-random challenge helpers: srandom(seed) + fieldElement::random() order */
void ogkr_seed(unsigned seed);
ofe ogkr_random_field(void);

/* Number of F elements in a full transcript of this circuit (incl. the trailing input-MLE). */
size_t ogkr_transcript_len(const ogkr_circuit* c);
/* Number of challenges the verifier draws. */
size_t ogkr_challenge_count(const ogkr_circuit* c);

/* Run the prover against the verifier's challenge order (seed as in F::init(): 3396).
 * transcript[] receives the prover messages; challenges[] (may be NULL) the drawn challenges.
 * Returns 0, or -1 if an assert gate is violated. prove_seconds (may be NULL): time spent inside
 * prover methods (the reference's `Prove Time`). */
int ogkr_prove(const ogkr_circuit* c, unsigned seed, ofe* transcript, ofe* challenges, double* prove_seconds);

/* Verify a transcript (verifier.cpp:134-337 checks + final input-MLE equality).
 * Returns 1 = accept, 0 = reject; fail_code/fail_layer describe the first failing check:
 * 1 phase1 round, 2 phase2 round, 3 semi-final (getFinalValue), 4 Liu round, 5 Liu final, 6 input. */
int ogkr_verify(const ogkr_circuit* c, unsigned seed, const ofe* transcript, int* fail_code, int* fail_layer);

/* Circuit evaluation only (prover.cpp:27-91): values[] sized sum over layers of layer_size. */
void ogkr_evaluate(const ogkr_circuit* c, ofe* values);

/* Stand-alone multilinear sumcheck over three tables of n = 2^log_n entries (config C2):
 * out[3*round + {0,1,2}] = (a,b,c) of each round, out[3*log_n .. +3) = final V, add, mult folded
 * values. r[round] = challenge bound after `round` (r[0] is consumed by round 1's successor). */
void ogkr_sumcheck_tables(const ofe* V, const ofe* add, const ofe* mult, int log_n, const ofe* r, ofe* out);

/* eq table: beta[i] = init * prod_k (i_k ? r_k : 1-r_k)   (utils.cpp:8-45) */
void ogkr_beta_table(ofe* beta, int n_bits, const ofe* r, ofe init);

#ifdef __cplusplus
}
#endif
#endif
