/* virgo_b200.h -- C ABI of the B200-native Virgo++ GKR sumcheck prover.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types. The C++ class
 * `prover` shipped in virgo-plus_b200/host/prover.{h,cpp} (same public API as the reference's
 * src/prover.h:12-42) forwards 1:1 to these entry points; INTEGRATION.md shows the binding.
 * Citations are into /root/reference.
 *
 * Conventions: every function returns 0 on success or a negative vp_status; vp_last_error()
 * returns a description of the last failure on the calling thread. There is NO CPU fallback: all
 * prover entry points fail with VP_ERR_CUDA when no sm_100 device / driver is usable.
 * A context is single-threaded (like the reference prover: one prover per process).
 *
 * Limits (checked; violations return VP_ERR_CIRCUIT / VP_ERR_ARG / VP_ERR_CUDA with a message, never undefined behaviour):
 *   - 2 <= layers <= 120; layer_size * instances <= 2^31 per layer (table indices are 32 bit), hence at most 31 rounds
 *     per sumcheck phase (the whole-proof kernel takes the phase's <= 32 challenges through its parameters);
 *   - in-layer gate indices, lv and dadId entries fit 32 bits; every dadId entry must be < the source layer's size;
 *   - .pws files: variable ids dense and < 2^40 (main.cpp indexes a vector by id); malformed numbers reject the file;
 *   - sharded contexts: world in {1, 2, 4, 8}, one process per GPU of ONE node;
 *   - witness inputs are read like the reference reads them, as F((long long) x) (prover.cpp:30-36,
 *     fieldElement.cpp:24-27): 0 <= x < p as is, x < 0 as p + x; other values are reduced mod p (the reference keeps
 *     them non-canonical). The circuit-level setters (vp_circuit_set_inputs, .pws inputs) only accept 0 <= x < p;
 *   - challenges must be canonical (both components < p): vp_set_challenges / vp_prove / vp_round reject others.
 */
#ifndef VIRGO_B200_H
#define VIRGO_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* bit-identical to virgo::fieldElement {real, img} (lib/virgo/src/fieldElement.hpp:96-97) */
typedef struct { uint64_t re, im; } vp_F;

typedef struct vp_circuit vp_circuit; /* host-side layered circuit (+ instance count) */
typedef struct vp_ctx vp_ctx;         /* prover context: owns all device memory */

typedef enum {
    VP_OK = 0,
    VP_ERR_ARG = -1,      /* bad argument / call out of protocol order */
    VP_ERR_CIRCUIT = -2,  /* malformed circuit or .pws */
    VP_ERR_CUDA = -3,     /* CUDA / NCCL failure, or no usable device */
    VP_ERR_ASSERT = -4,   /* an is_assert gate evaluates to non-zero (prover.cpp:18-21, :211) */
    VP_ERR_NOMEM = -5
} vp_status;

const char* vp_last_error(void);
const char* vp_version(void);

/* ------------------------------------------------------------------ circuit (host side)
 * Replaces parse()+DAG_to_layered() (src/main.cpp:15-137,176-231) and layeredCircuit::subsetInit()
 * (src/circuit.cpp:43-80), including their observable quirks. */
int vp_circuit_load_pws(const char* path, vp_circuit** out);
int vp_circuit_load_pws_text(const char* text, size_t len, vp_circuit** out);
/* Synthetic random add/mul circuit, semantics of layeredCircuit::randomize (circuit.cpp:17-41). */
int vp_circuit_random(int n_layers, int log_size, uint64_t seed, vp_circuit** out);
/* Build from flat arrays (concatenated over layers 0..n-1, gate_off-style: layer i occupies
 * [sum_{j<i} layer_size[j], ...)). Layer-0 entries carry the input VALUE in u (like gate.u of an
 * Input gate, main.cpp:108-110). If dad_size == NULL the subsets and lv are derived with the
 * subsetInit rule; otherwise lv / dad_size[i*n+l] / dad_id (concatenated in (i,l) order) are taken
 * as given (what the reference's `layeredCircuit` holds after c.subsetInit()). c / is_assert may be
 * NULL. */
int vp_circuit_from_arrays(int n_layers, const uint64_t* layer_size, const uint8_t* ty, const int32_t* l,
                           const uint64_t* u, const uint64_t* v, const uint64_t* lv, const vp_F* c,
                           const uint8_t* is_assert, const uint64_t* dad_size, const uint64_t* dad_id,
                           vp_circuit** out);
/* K data-parallel instances of the same template (instance-major replication);
 * inputs are drawn like the reference draws them at parse time (main.cpp:188). */
int vp_circuit_replicate(const vp_circuit* c, uint64_t instances, vp_circuit** out);
/* Materialise the instances into one flat circuit (instances == 1), subsets re-derived. */
int vp_circuit_expand(const vp_circuit* c, vp_circuit** out);
void vp_circuit_free(vp_circuit* c);

int vp_circuit_num_layers(const vp_circuit* c);
uint64_t vp_circuit_instances(const vp_circuit* c);
uint64_t vp_circuit_layer_size(const vp_circuit* c, int layer);      /* of ONE instance */
int vp_circuit_bit_length(const vp_circuit* c, int layer);          /* of the replicated layer */
uint64_t vp_circuit_dad_size(const vp_circuit* c, int layer, int src); /* of ONE instance */
int vp_circuit_max_dad_bit_length(const vp_circuit* c, int layer);  /* -1: phase 2 skipped */
uint64_t vp_circuit_total_gates(const vp_circuit* c);               /* non-input gates, all instances */
uint64_t vp_circuit_num_inputs(const vp_circuit* c);                /* instances * layer_size(0) */
/* Export the gates of one template layer (arrays of layer_size entries; any pointer may be NULL). */
int vp_circuit_export_layer(const vp_circuit* c, int layer, uint8_t* ty, int32_t* l, uint32_t* u, uint32_t* v,
                            uint32_t* lv, vp_F* cst, uint8_t* is_assert);
int vp_circuit_export_dad(const vp_circuit* c, int layer, int src, uint32_t* dad_id /* dad_size entries */);
int vp_circuit_get_inputs(const vp_circuit* c, uint64_t* out /* num_inputs */);
int vp_circuit_set_inputs(vp_circuit* c, const uint64_t* in /* num_inputs, each < p */);

/* Challenges in the verifier's draw order (src/verifier.cpp:144-145,196,202,236,278-279) from glibc
 * random() after srand(seed) (fieldElement.cpp:106-124); seed 3396 reproduces F::init(). */
size_t vp_challenge_count(const vp_circuit* c);
int vp_draw_challenges(const vp_circuit* c, unsigned seed, vp_F* out /* vp_challenge_count */);
/* The same generator as a plain stream: the first n values of fieldElement::random() after srand(seed). */
int vp_draw_field(unsigned seed, size_t n, vp_F* out);
/* Prover messages: Vres; per layer top..1: bl(i-1) x (a,b,c), claim_u, [maxDad(i) x (a,b,c), claims_v[0..i)],
 * bl(i-1) x (a,b,c), claim_liu; finally the input-layer MLE. */
size_t vp_transcript_len(const vp_circuit* c);

/* Transcript containers. GKRProof byte stream = the layout of GKRProof::write (src/GKRProof.hpp:23-58; the struct is
 * dead code in the reference): members final_claims_u, final_claims, final_claims_v, polys_u, polys_v,
 * polys, outer vectors indexed by layer id; the reference's undefined poly_proof part is replaced by a trailer
 * {u64 2, Vres, input MLE}. out == NULL: only *len is set. Text dump: "TAG real img" lines. */
int vp_transcript_to_gkrproof(const vp_circuit* c, const vp_F* transcript, unsigned char* out, size_t cap, size_t* len);
int vp_gkrproof_to_transcript(const vp_circuit* c, const unsigned char* bytes, size_t len, vp_F* transcript);
int vp_transcript_text(const vp_circuit* c, const vp_F* transcript, const vp_F* challenges, char* out, size_t cap,
                       size_t* len);

/* ------------------------------------------------------------------ prover
 * One context = one `prover` object (src/prover.h:12-67) on one GPU. */
int vp_create(const vp_circuit* c, int device, vp_ctx** out);
/* Sharded context: this process is `rank` of `world` GPUs of one box; nccl_id is the 128-byte
 * ncclUniqueId shared by all ranks (vp_nccl_unique_id on rank 0, broadcast by the caller). */
int vp_nccl_unique_id(uint8_t out[128]);
int vp_create_sharded(const vp_circuit* c, int device, int rank, int world, const uint8_t nccl_id[128],
                      vp_ctx** out);
void vp_destroy(vp_ctx* ctx);
/* Host-only: how phase (1, 2 = phase 2, 3 = Liu) of `layer` is dealt out to `rank` of `world` GPUs.
 * out: 10 values per table {bits, live, sharded, m, row_lo, row_hi, local_len, present, n_blocks, reversed}:
 * the rank holds the live table entries [row_lo, row_hi). */
int vp_shard_describe(const vp_circuit* c, int world, int rank, int layer, int phase, uint32_t* out, size_t cap,
                      size_t* n_tables);
/* Host-only: index map of a rank holding the table entries [lo, hi): returns 1 and *local if idx belongs to it. */
/* Host-only (no device needed): the instances [lo[l], hi[l]) of layer l that `rank` of a `world`-GPU sharded context
   evaluates -- the hull of what its tables read from layer l and from every layer above it. lo, hi: n_layers entries. */
int vp_shard_eval_ranges(const vp_circuit* c, int world, int rank, uint32_t* lo, uint32_t* hi);
int vp_shard_map_index(uint32_t lo, uint32_t hi, uint32_t idx, uint32_t* local);

/* SHARDED CONTEXTS and the method-by-method API. Every entry point below except vp_get_values also works on a sharded
 * context, as a COLLECTIVE call: all ranks call it in the same order with the same arguments and all get the same
 * results (like the reference's single prover, whose messages every rank's verifier copy sees). A round of a sharded
 * phase folds the rank's own table rows and exchanges 48 bytes per rank (the partial round polynomial, partial add_term
 * included); after the local rounds the remaining <= 2^10 values per table are gathered once and the last rounds run
 * replicated (src/prover.cpp:436-455 per round). */
/* Upload the witness inputs (instances * layer_size(0) values < p); default: the circuit's own. */
int vp_set_inputs(vp_ctx* ctx, const uint64_t* inputs, size_t n);
/* prover::evaluate (prover.cpp:27-91) + the assert check of the constructor (:16-24). */
int vp_evaluate(vp_ctx* ctx);
/* Copy circuitValue[layer] to the host (layer 0 is what commit_private hands to the PC, :524-530). */
int vp_get_values(vp_ctx* ctx, int layer, vp_F* out, size_t n);

/* prover::Vres (prover.cpp:99-129): MLE of the output layer at r[0..n). */
int vp_vres(vp_ctx* ctx, const vp_F* r, int n, vp_F* out);
/* prover::sumcheckInitAll (:162-170) */
int vp_sumcheck_init_all(vp_ctx* ctx, const vp_F* r_last, int n);
/* prover::sumcheckInit (:177-184): step to the next layer (top -> 1). */
int vp_sumcheck_init(vp_ctx* ctx);
/* prover::sumcheckInitPhase1 (:189-280) / Phase2 (:282-367) / Liu (:369-420; sig has n entries,
 * n >= n_layers - layer + 1). */
int vp_init_phase1(vp_ctx* ctx, const vp_F* assert_random);
int vp_init_phase2(vp_ctx* ctx);
int vp_init_liu(vp_ctx* ctx, const vp_F* sig, int n);
/* prover::sumcheckUpdatePhase1 / Phase2 / LiuUpdate (:422-455): phase = 1, 2, 3. out_abc = the
 * round's quadratic_poly {a,b,c}. */
int vp_round(vp_ctx* ctx, int phase, const vp_F* previous_random, vp_F out_abc[3]);
/* prover::sumcheckFinalize1 (:494-501), Finalize2 (:504-516; claims has `layer` entries),
 * LiuFinalize (:518-521). */
int vp_finalize1(vp_ctx* ctx, const vp_F* previous_random, vp_F* claim);
int vp_finalize2(vp_ctx* ctx, const vp_F* previous_random, vp_F* claims, int n);
int vp_finalize_liu(vp_ctx* ctx, const vp_F* previous_random, vp_F* claim);
/* prover::inner_prod(circuitValue[0], pub, n) as used by commit_public (:532-546). */
int vp_inner_prod(vp_ctx* ctx, const vp_F* pub, size_t n, vp_F* out);
/* prover::inner_prod(a, b, n) for two HOST vectors (the public method, prover.h:40). */
int vp_dot_host(vp_ctx* ctx, const vp_F* a, const vp_F* b, size_t n, vp_F* out);
/* Input-layer MLE at r[0..n): <circuitValue[0], eq(r,.)> without materialising eq on the host. */
int vp_input_mle(vp_ctx* ctx, const vp_F* r, int n, vp_F* out);
/* prover::proofSize() in bytes (:451,:500,:512) and the accumulated device time inside prover
 * entry points in seconds (the analogue of proveTime(), :549-551). */
uint64_t vp_proof_size_bytes(const vp_ctx* ctx);
double vp_prove_seconds(const vp_ctx* ctx);

/* ------------------------------------------------------------------ whole proof, no per-round host sync
 * The reference verifier's challenges do not depend on prover messages, so the whole
 * stream can be handed over up front. vp_prove runs evaluate + every phase of every layer on the
 * device and returns the transcript (vp_transcript_len entries).
 * host_io != 0: `inputs`/`challenges`/`transcript` are HOST buffers (copied inside the call).
 * host_io == 0: inputs already uploaded with vp_set_inputs and challenges with vp_set_challenges;
 *               the transcript stays on the device until vp_get_transcript. */
int vp_set_challenges(vp_ctx* ctx, const vp_F* challenges, size_t n);
int vp_prove(vp_ctx* ctx, int host_io, const uint64_t* inputs, size_t n_inputs, const vp_F* challenges,
             size_t n_challenges, vp_F* transcript, size_t transcript_cap);
/* Sharded contexts: a rank only uploads the witness of the data-parallel instances [first, end) it evaluates (its own
 * K/world slice plus the few neighbours its tables read, vp_shard_eval_ranges). vp_prove_local is vp_prove(host_io=1)
 * for a caller that holds just that slice: local_inputs = (end - first) * layer_size(0) values, instance-major.
 * On an unsharded context the range is [0, instances) and the two calls are the same. */
int vp_input_range(const vp_ctx* ctx, uint64_t* first_instance, uint64_t* end_instance);
int vp_prove_local(vp_ctx* ctx, const uint64_t* local_inputs, size_t n_local, const vp_F* challenges, size_t n_challenges,
                   vp_F* transcript, size_t transcript_cap);
/* ------------------------------------------------------------------ polynomial commitment, commit phase
 * Replaces poly_commit_prover::commit_private_array (lib/virgo/src/poly_commit.h:41-124: per-slice inverse FFT + 32x
 * Reed-Solomon extension, RS_polynomial.cpp:26-220), fri::request_init_commit (fri.cpp:36-139: SHA3-256 leaf chains,
 * my_hhash.h:27-33) and merkle_tree_prover::create_tree (merkle_tree.cpp:7-51) as prover::commit_private uses them
 * (prover.cpp:524-530: circuitValue[0] padded to 2^bitLength, ONE zero mask element -- other masks return VP_ERR_ARG).
 * root = the 32-byte Merkle root (the __hhash_digest commit_private returns). The codeword array l_eval
 * [65 * slice_size], the leaf hashes [slice_size / 2 * 32 B] and the Merkle tree [slice_size * 32 B, array heap, node 1 =
 * root] stay on the device; vp_commit_export copies them out for the reference's CPU opening phase (any pointer may be
 * NULL). slice_size = 2^(bitLength(0) - 1). Needs bitLength(0) >= 6; unsharded contexts.
 * Not reproduced: with bitLength(0) == 8 the reference's 4-point inverse FFT returns uninitialised scratch memory
 * (RS_polynomial.cpp:100: `blk_size / packed_size * 2` iterations == 0); this library computes the transform. */
int vp_commit_private(vp_ctx* ctx, const vp_F* mask, size_t n_mask, uint8_t root[32]);
int vp_commit_export(vp_ctx* ctx, vp_F* l_eval, uint8_t* leaf_hash, uint8_t* tree);
uint64_t vp_commit_slice_size(const vp_ctx* ctx);
float vp_last_commit_ms(const vp_ctx* ctx);   /* device time of the last vp_commit_private */
/* prover::commit_public (prover.cpp:542-546) -> commit_public_array (poly_commit.h:126-349) after vp_commit_private: pub =
 * the public array (the verifier's eq table over r_liu, verifier.cpp:367-381; n <= 2^bitLength(0) canonical elements), mask
 * = the public mask (one zero: others return VP_ERR_ARG). Encodes pub like the private array, gets the 2n coefficients
 * of l*q per slice from its values on the 2n-th roots, extends the quotient h (l q = g + (x^n - 1) h) to all points, builds
 * the virtual oracle (g - const) n / x and all_sum[65], and commits to h_eval_arr (second Merkle tree): root_h.
 * vp_commit_public_export: h_eval_arr [65 * slice_size], virtual_oracle_witness [64 * slice_size] (interleaved like
 * fri::virtual_oracle_witness), leaf hashes and tree of the second commitment (any pointer may be NULL).
 * Not reproduced: bitLength(0) == 7 (the reference's 4-point inverse FFT, see above). */
int vp_commit_public(vp_ctx* ctx, const vp_F* pub, size_t n, const vp_F* mask, size_t n_mask, uint8_t root_h[32], vp_F all_sum[65]);
int vp_commit_public_export(vp_ctx* ctx, vp_F* h_eval, vp_F* vow, uint8_t* leaf_hash, uint8_t* tree);
/* The 64 codewords of commitment `which` (0: after vp_commit_private, 1: after vp_commit_public) in the layout of
 * fri::witness_rs_codeword_interleaved[which] (fri.cpp:69-96): out[(j << 7) | (slice << 1) | h] = codeword[slice][j + h *
 * slice_size / 2], 64 * slice_size elements; transposed on the device. Before the first vp_fri_commit_steps call. */
int vp_commit_export_interleaved(vp_ctx* ctx, int which, vp_F* out);
/* The same on a host array of n canonical field elements, zero-padded to 2^log_len (6 <= log_len <= 30). */
int vp_pc_commit(int device, const vp_F* array, size_t n, int log_len, uint8_t root[32], vp_F* l_eval, uint8_t* leaf_hash,
                 uint8_t* tree, float* device_ms);
int vp_pc_commit_public(int device, const vp_F* array, size_t n, const vp_F* pub, size_t n_pub, int log_len, uint8_t root_l[32],
                        uint8_t root_h[32], vp_F all_sum[65], vp_F* h_eval, vp_F* vow, float* device_ms);

/* FRI commit phase (fri::commit_phase_step, lib/virgo/src/fri.cpp:289-418, as poly_commit_prover::commit_phase drives it,
 * vpd_verifier.cpp:43-73) on the virtual oracle vp_commit_public left on the device. One step per fold challenge: every
 * slice's codeword f on the M-th roots becomes g(x^2) = (f(x) + f(-x))/2 + r (f(x) - f(-x))/(2x) on the M/2-th roots, the
 * M/4 leaves (one pair of opposite points of all 64 slices + the zero mask pair each) are hashed and the level's tree
 * built; roots[k] = the root after step k. vp_fri_steps(ctx) = bitLength(0) - 6 steps leave 32 points per slice (the
 * reference stops there). vp_fri_export_level: level lvl's codewords in the layout of fri::cpd.rs_codeword[lvl]
 * (64 * (slice_size >> (lvl+1)) elements) and its tree (array heap, (slice_size >> (lvl+1)) nodes of 32 bytes); any pointer
 * may be NULL. vp_fri_restart: back to step 0 on the same virtual oracle. Challenges must be canonical. */
int vp_fri_commit_steps(vp_ctx* ctx, const vp_F* randomness, int n_steps, uint8_t* roots /* n_steps * 32 */);
int vp_fri_steps(const vp_ctx* ctx);
int vp_fri_restart(vp_ctx* ctx);
int vp_fri_export_level(vp_ctx* ctx, int lvl, vp_F* rs_codeword, uint8_t* merkle);
/* Stand-alone: both commitments of host arrays (as vp_pc_commit_public), then n_steps <= log_len - 6 FRI steps.
 * codes / trees (may be NULL): the levels back to back. 7 <= log_len <= 30. */
int vp_pc_fri(int device, const vp_F* array, size_t n, const vp_F* pub, size_t n_pub, int log_len, const vp_F* randomness, int n_steps,
              uint8_t root_l[32], uint8_t root_h[32], uint8_t* roots, vp_F* codes, uint8_t* trees, float* device_ms);

/* ------------------------------------------------------------------ the commitment's inner GKR
 * fft_circuit_gkr::fft_gkr (lib/virgo/src/fft_circuit_GKR.cpp:833-849), which poly_commit_verifier::verify_poly_commitment
 * runs once per opening (vpd_verifier.cpp:92): prover and verifier of a layered GKR over a fixed circuit family -- the eq
 * table of a random point, lg_size inverse-FFT butterfly layers, a scaling, 64 x 2^lg_size products with the powers of 64
 * random points and their row sums -- walked from the outputs back to the eq table with 2 + 2 lg_size sumchecks. The
 * reference draws all randomness with fieldElement::random() and none of it depends on a prover message; here the caller
 * passes it as one array in the reference's draw order:
 *   r[lg] | x[64] | r_0[lg+10] | r_1[lg+10] | addition layer: r_u[lg+6], r_v[lg+6] | mult layer: r_u[lg], r_v[lg] |
 *   per butterfly layer (lg of them): r_u[lg], r_v[lg], alpha, beta          (vp_fft_gkr_rnd_count(lg) elements, canonical)
 * The prover side (circuit evaluation, all sumcheck tables and rounds) runs on the device without a host round trip; the
 * verifier's claim chain and closed-form wiring predicates run on the host afterwards.
 * Outputs: proof_size = the reference's `ps` (48 bytes per round + extension_gkr's count), ok = 1 if every verifier check
 * passed (the reference prints "Error, fft gkr failed" otherwise), verifier_seconds / prover_seconds = its `vt` / `pt`;
 * optional (NULL to skip): layers = E, F_{lg-1}..F_0, S (2^lg each), P (64 * 2^lg), O (64); polys = 3 elements per round in
 * protocol order (vp_fft_gkr_poly_count(lg) rounds); claims = the running claim a_0, after the addition / mult /
 * intermediate layer, after every butterfly layer, then the final alpha, beta (4 + lg + 2 elements). 1 <= lg_size <= 24. */
/* The two sumcheck objects of a size (lg_size and lg_size + 6 variables: device tables of 3 x 2 x 16 bytes per entry) are kept
 * per (device, lg_size) for the next call; vp_fft_gkr_release() frees them. Calls are serialised by an internal lock. */
void vp_fft_gkr_release(void);
size_t vp_fft_gkr_rnd_count(int lg_size);
size_t vp_fft_gkr_poly_count(int lg_size);
int vp_fft_gkr(int device, int lg_size, const vp_F* rnd, size_t n_rnd, vp_F* layers, vp_F* polys, size_t polys_cap, vp_F* claims,
               int* proof_size, int* ok, double* verifier_seconds, double* prover_seconds, float* device_ms);

/* ------------------------------------------------------------------ Fiat-Shamir mode
 * The reference ships a hash-based challenge source, transcriptCache (lib/virgo/src/transcriptCache.hpp:14-50: a byte
 * pool hashed with SHA3-256 per challenge, challenge = first two digest words mod p), but never calls it. This mode uses
 * exactly that class, fed with seed[32] and then every prover message (16 bytes {u64 real, u64 img} each) in emission
 * order; a round's challenge is drawn right after the round's polynomial was stored (order: host/fiat_shamir.h). Only the
 * one-round-per-launch path applies (a challenge depends on the previous message). vp_prove_fs uses the resident inputs
 * and returns the transcript and, optionally, the challenges it drew (usual layout, unused slots zero);
 * vp_fs_challenges recomputes them from a transcript alone; vp_verify_fs = vp_fs_challenges + vp_verify (fail_code 7:
 * a non-canonical transcript element). Not parity-bound to the reference (it has no such mode); pinned against the
 * oracle's restatement of transcriptCache. */
int vp_prove_fs(vp_ctx* ctx, const uint8_t seed[32], vp_F* transcript, size_t transcript_cap, vp_F* challenges, size_t challenges_cap);
int vp_fs_challenges(const vp_circuit* c, const uint8_t seed[32], const vp_F* transcript, size_t n, vp_F* challenges, size_t cap);
int vp_verify_fs(vp_ctx* ctx, const uint8_t seed[32], const vp_F* transcript, size_t n, int* accept, int* fail_code, int* fail_layer);

/* Self-test of the device-only arithmetic paths of csrc/field.cuh (inline-PTX carry chains, mul.wide / mad.wide):
   runs one routine on n caller-provided operand triples and returns the results, so that tests can feed edge values
   (0, 1, p-1, p, limb boundaries) and compare with big-integer arithmetic.  op: 0 weak fold a + c*(b-a); 1 the same for
   base-field a, b; 2 lazy dot product sum a_i*b_i -> out[0]; 3 {fp_reduce_ut_weak(a.re, a.im, b.re), fp_weak(b.im)};
   4 c + a*b (weak); 5 c + a*b.re (weak); 6 lazy sum (b_i - a_i)*c_i.re -> out[0]; 7 like 0 with the folded difference. */
int vp_selftest_field(int device, int op, const vp_F* a, const vp_F* b, const vp_F* c, vp_F* out, size_t n);

/* Verifier: replaces verifier::verify's checks, src/verifier.cpp:134-337, on a finished transcript:
   the O(#gates) sums of predicatePhase1/2 (:63-113) and verifyLiu's gr (:311-323) run on the device, the round
   checks / getFinalValue (:115-132) / Liu check on the host. Uses the context's circuit, its resident inputs
   (vp_set_inputs or vp_prove with host_io) and challenge stream (vp_set_challenges / vp_prove). *accept = 1, or 0 with
   *fail_code = 1 phase-1 round, 2 phase-2 round, 3 layer final value, 4 Liu round, 5 Liu final, 6 input layer, and
   *fail_layer (either may be NULL). The polynomial-commitment opening (verifyPoly) is out of scope: the input-layer
   MLE is recomputed from the inputs. On a sharded context this is a COLLECTIVE call: every rank passes the same
   transcript, sums the gates of its own slice of the instances, the partial sums meet in one all-gather and every
   rank returns the same verdict. */
int vp_verify(vp_ctx* ctx, const vp_F* transcript, size_t n, int* accept, int* fail_code, int* fail_layer);
int vp_get_transcript(vp_ctx* ctx, vp_F* transcript, size_t cap);
/* Device-side time of the last vp_prove (CUDA events on the context's stream), milliseconds. */
float vp_last_prove_ms(const vp_ctx* ctx);
/* Number of kernel launches issued by the last vp_prove. */
uint64_t vp_last_prove_launches(const vp_ctx* ctx);
/* Run every kernel of the context on the caller's stream (cudaStream_t; NULL: back to an own stream). */
int vp_set_stream(vp_ctx* ctx, void* cuda_stream);
/* Per-kernel-class device timing (CUDA events around every launch while on). Classes, in order:
 * 0 round(fold) [K6], 1 round(first), 2 phase-1 init [K3], 3 phase-2 init [K4], 4 Liu init [K5],
 * 5 evaluate [K1], 6 other. bytes = algorithmic bytes of the launches. */
int vp_set_profiling(vp_ctx* ctx, int on);
/* Whole-proof mode runs phase 1, phase 2 and the Liu phases of vp_prove on up to three concurrent streams ("lanes",
   DESIGN.md). lanes = 1 runs them one after the other (used by bench.py's instrumented pass so that a launch's
   CUDA-event duration is its own), 2 = Liu phases on their own lane, 3 = one lane per phase kind, 6 = two sets of
   three lanes taking alternate layers (the default on unsharded contexts). Returns the setting in force. */
int vp_set_lanes(vp_ctx* ctx, int lanes);
int vp_get_profile(vp_ctx* ctx, double* ms, double* bytes, uint64_t* launches, int n_classes);
/* The context's CUDA stream (cudaStream_t) so callers can bracket it with their own events. */
void* vp_stream(vp_ctx* ctx);

/* ------------------------------------------------------------------ stand-alone sumcheck (config C2)
 * Three tables V, add, mult of 2^log_n entries (device-resident copies are made once); runs the
 * log_n rounds of sumcheckUpdateEach with challenges r[0..log_n) and returns 3*log_n + 3 values:
 * per round (a,b,c), then the three fully-folded table values. */
typedef struct vp_sumcheck vp_sumcheck;
int vp_sumcheck_create(int log_n, int device, vp_sumcheck** out);
int vp_sumcheck_load(vp_sumcheck* s, const vp_F* V, const vp_F* add, const vp_F* mult); /* host -> device */
int vp_sumcheck_fill_random(vp_sumcheck* s, uint64_t seed);  /* SplitMix64 per entry, on device */
int vp_sumcheck_export(vp_sumcheck* s, vp_F* V, vp_F* add, vp_F* mult);  /* device -> host (pristine copy) */
int vp_sumcheck_run(vp_sumcheck* s, const vp_F* r, vp_F* out /* 3*log_n+3 */, float* device_ms);
/* Same outputs, all rounds in one cooperative launch, two rounds per pass (the challenges are all known). */
int vp_sumcheck_run_fused(vp_sumcheck* s, const vp_F* r, vp_F* out, float* device_ms);
/* profiling aid: ns time stamps of block 0 at {pass start, work done, barrier passed, pass end} per pass of the last fused run */
int vp_sumcheck_pass_stamps(vp_sumcheck* s, unsigned long long* out, int n);
/* per-round device time of the last vp_sumcheck_run (log_n floats, ms) */
int vp_sumcheck_round_ms(vp_sumcheck* s, float* out);
void vp_sumcheck_destroy(vp_sumcheck* s);

#ifdef __cplusplus
}
#endif
#endif
