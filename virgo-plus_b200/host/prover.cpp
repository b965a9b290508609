// Drop-in `prover` over the B200 engine; see prover.h. Method-by-method counterpart of
// /root/reference/src/prover.cpp (line ranges cited per method).
#include "prover.h"

#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "virgo_b200.h"

static_assert(sizeof(F) == sizeof(vp_F), "fieldElement must be {u64 real; u64 img;}");

namespace {
[[noreturn]] void die(const char *what) {
    fprintf(stderr, "virgo_b200 prover: %s failed: %s\n", what, vp_last_error());
    exit(EXIT_FAILURE);
}
inline void ck(int rc, const char *what) {
    if (rc != 0) die(what);
}
inline const vp_F *cf(const F *p) { return reinterpret_cast<const vp_F *>(p); }
inline vp_F *mf(F *p) { return reinterpret_cast<vp_F *>(p); }
}  // namespace

// prover.cpp:14-25: evaluate, then fail if an assert gate is non-zero.
prover::prover(const layeredCircuit &cir) : C(cir), circ(nullptr), ctx(nullptr), sumcheckLayerId(0), world(1), gpu_commit(false) {
    const int n = C.size;
    std::vector<uint64_t> layer_size(n), u, v, lv, dad_size((size_t)n * n, 0), dad_id;
    std::vector<uint8_t> ty, is_assert;
    std::vector<int32_t> l;
    std::vector<vp_F> cst;
    for (int i = 0; i < n; ++i) {
        const layer &L = C.circuit[i];
        layer_size[i] = L.size;
        for (u64 g = 0; g < L.size; ++g) {
            const gate &G = L.gates[g];
            ty.push_back((uint8_t)G.ty);
            l.push_back(G.l);
            u.push_back(G.u);
            v.push_back(G.v);
            lv.push_back(G.lv);
            cst.push_back(vp_F{G.c.real, G.c.img});
            is_assert.push_back(G.is_assert ? 1 : 0);
        }
        for (int s = 0; s < i; ++s) {
            dad_size[(size_t)i * n + s] = L.dadSize[s];
            for (u64 x = 0; x < L.dadSize[s]; ++x) dad_id.push_back(L.dadId[s][x]);
        }
    }
    if (dad_id.empty()) dad_id.push_back(0);
    ck(vp_circuit_from_arrays(n, layer_size.data(), ty.data(), l.data(), u.data(), v.data(), lv.data(), cst.data(),
                              is_assert.data(), dad_size.data(), dad_id.data(), &circ),
       "vp_circuit_from_arrays");
    // One GPU: VP_DEVICE (default 0). Several GPUs of one box: start VP_WORLD copies of the program, one per GPU, with
    // VP_RANK = 0 .. VP_WORLD-1 (and VP_NCCL_ID_FILE = a path all of them can reach): the copies run the same verifier
    // (its challenges are deterministic), every prover call is a collective of the sharded engine.
    const char *dev = getenv("VP_DEVICE"), *w = getenv("VP_WORLD");
    world = w ? atoi(w) : 1;
    if (world > 1) {
        const char *r = getenv("VP_RANK"), *idf = getenv("VP_NCCL_ID_FILE");
        if (!r || !idf) { fprintf(stderr, "virgo_b200 prover: VP_WORLD needs VP_RANK and VP_NCCL_ID_FILE\n"); exit(EXIT_FAILURE); }
        const int rank = atoi(r);
        uint8_t id[128];
        if (rank == 0) {   // publish the NCCL id: write to a temporary name, then rename (readers never see a partial file)
            ck(vp_nccl_unique_id(id), "vp_nccl_unique_id");
            const std::string tmp = std::string(idf) + ".tmp";
            FILE *f = fopen(tmp.c_str(), "wb");
            if (!f || fwrite(id, 1, 128, f) != 128) { fprintf(stderr, "virgo_b200 prover: cannot write %s\n", tmp.c_str()); exit(EXIT_FAILURE); }
            fclose(f);
            rename(tmp.c_str(), idf);
        } else {
            for (int tries = 0;; ++tries) {
                FILE *f = fopen(idf, "rb");
                if (f) {
                    const size_t got = fread(id, 1, 128, f);
                    fclose(f);
                    if (got == 128) break;
                }
                if (tries > 3000) { fprintf(stderr, "virgo_b200 prover: no NCCL id in %s\n", idf); exit(EXIT_FAILURE); }
                usleep(10000);
            }
        }
        ck(vp_create_sharded(circ, dev ? atoi(dev) : rank, rank, world, id, &ctx), "vp_create_sharded");
    } else ck(vp_create(circ, dev ? atoi(dev) : 0, &ctx), "vp_create");
    evaluate();
}

prover::~prover() {
    vp_destroy(ctx);
    vp_circuit_free(circ);
}

// prover.cpp:27-91
void prover::evaluate() {
    int rc = vp_evaluate(ctx);
    if (rc == VP_ERR_ASSERT) {
        fprintf(stderr, "FAIL ON: %s\n", vp_last_error());
        exit(EXIT_FAILURE);
    }
    ck(rc, "vp_evaluate");
}

// prover.cpp:131-155: only sizes host-side state in the reference; the engine sized everything at create.
void prover::init() {}

// prover.cpp:99-129
F prover::Vres(const vector<F>::const_iterator &r_0, int r_0_size) {
    F out;
    ck(vp_vres(ctx, cf(&*r_0), r_0_size, mf(&out)), "vp_vres");
    return out;
}

// prover.cpp:162-170
void prover::sumcheckInitAll(const vector<F>::const_iterator &r_last) {
    sumcheckLayerId = C.size;
    ck(vp_sumcheck_init_all(ctx, cf(&*r_last), C.circuit[C.size - 1].bitLength), "vp_sumcheck_init_all");
}

// prover.cpp:177-184
void prover::sumcheckInit() {
    --sumcheckLayerId;
    ck(vp_sumcheck_init(ctx), "vp_sumcheck_init");
}

// prover.cpp:189-280 (incl. its progress lines on stderr)
void prover::sumcheckInitPhase1(const F &assert_random) {
    fprintf(stderr, "sumcheck level %d, phase1 init start\n", sumcheckLayerId);
    ck(vp_init_phase1(ctx, cf(&assert_random)), "vp_init_phase1");
    fprintf(stderr, "sumcheck level %d, phase1 init finished\n", sumcheckLayerId);
}

// prover.cpp:282-367
void prover::sumcheckInitPhase2() {
    fprintf(stderr, "sumcheck level %d, phase2 init start\n", sumcheckLayerId);
    ck(vp_init_phase2(ctx), "vp_init_phase2");
}

// prover.cpp:369-420: reads s[0 .. C.size - layer + 1) of the verifier's sigma vector
void prover::sumcheckInitLiu(vector<F>::const_iterator s) {
    ck(vp_init_liu(ctx, cf(&*s), C.size - sumcheckLayerId + 1), "vp_init_liu");
}

quadratic_poly prover::round(int phase, const F &previousRandom) {
    F abc[3];
    ck(vp_round(ctx, phase, cf(&previousRandom), mf(abc)), "vp_round");
    return quadratic_poly(abc[0], abc[1], abc[2]);
}
// prover.cpp:422-434
quadratic_poly prover::sumcheckUpdatePhase1(const F &previousRandom) { return round(1, previousRandom); }
quadratic_poly prover::sumcheckUpdatePhase2(const F &previousRandom) { return round(2, previousRandom); }
quadratic_poly prover::sumcheckLiuUpdate(const F &previousRandom) { return round(3, previousRandom); }

// prover.cpp:494-501
void prover::sumcheckFinalize1(const F &previousRandom, F &claim) {
    ck(vp_finalize1(ctx, cf(&previousRandom), mf(&claim)), "vp_finalize1");
}
// prover.cpp:504-516: writes claims[0 .. layer)
void prover::sumcheckFinalize2(const F &previousRandom, vector<F>::iterator claims) {
    ck(vp_finalize2(ctx, cf(&previousRandom), mf(&*claims), sumcheckLayerId), "vp_finalize2");
}
// prover.cpp:518-521
void prover::sumcheckLiuFinalize(const F &previousRandom, F &claim) {
    ck(vp_finalize_liu(ctx, cf(&previousRandom), mf(&claim)), "vp_finalize_liu");
}

// prover.cpp:549-555
double prover::proveTime() const { return vp_prove_seconds(ctx); }
double prover::proofSize() const { return (double)vp_proof_size_bytes(ctx) / 1024.0; }

#ifdef USE_VIRGO
// prover.cpp:524-530 -> poly_commit_prover::commit_private_array (lib/virgo/src/poly_commit.h:41-124). The heavy part --
// 64 inverse FFTs, the 32x Reed-Solomon extension, 65 SHA3 evaluations per leaf and the Merkle tree -- runs on the
// device (vp_commit_private). The reference's OPENING phase (commit_public_array, the FRI rounds, the verifier's
// Merkle checks) stays its own CPU code and reads process-global arrays that commit_private_array and
// fri::request_init_commit (fri.cpp:36-139) leave behind: those are allocated here exactly as there and filled from the
// device results. VP_CPU_COMMIT=1 keeps the reference's CPU commit instead (A/B timing).
namespace virgo {
extern int witness_merkle_size[2];   // fri.cpp:22 (not declared in fri.h)
namespace fri {
// the reference's own CPU step, kept under this name when fri.cpp is compiled with
// -Dcommit_phase_step=commit_phase_step_reference
__hhash_digest commit_phase_step_reference(fieldElement r);
}
namespace fft_circuit_gkr {
int fft_gkr(int lg_size, double &vt, int &ps, double &pt);             // fft_circuit_GKR.h:4
int fft_gkr_reference(int lg_size, double &vt, int &ps, double &pt);   // the reference's own, renamed the same way
}
}
// the context whose device holds the virtual oracle (set by commit_public): fri::commit_phase_step and
// fft_circuit_gkr::fft_gkr are free functions
static vp_ctx *g_fri_ctx = nullptr;
static int g_device = 0;
virgo::__hhash_digest prover::commit_private() {
    using namespace virgo;
    std::vector<F> mask(1, F_ZERO);
    const int bl = C.circuit[0].bitLength;
    if (getenv("VP_CPU_COMMIT") || bl < 6 || world > 1) {   // (the device commit needs the whole input layer on one GPU)
        input_values.assign(1ULL << bl, F_ZERO);
        for (u64 g = 0; g < C.circuit[0].size; ++g) input_values[g] = F((long long)C.circuit[0].gates[g].u);   // prover.cpp:30-36
        return poly_prover.commit_private_array(input_values.data(), bl, mask);
    }
    gpu_commit = true;
    const auto t0 = std::chrono::high_resolution_clock::now();
    __hhash_digest root;
    double lap_ms[4] = {0, 0, 0, 0};
    auto t_lap = t0;
    auto lap = [&](int k) {
        const auto now = std::chrono::high_resolution_clock::now();
        lap_ms[k] = std::chrono::duration<double>(now - t_lap).count() * 1e3;
        t_lap = now;
    };
    ck(vp_commit_private(ctx, cf(mask.data()), mask.size(), reinterpret_cast<uint8_t *>(&root)), "vp_commit_private");
    lap(0);
    // ---- poly_commit.h:45-67: slicing parameters, the (padded) mask
    poly_commit::pre_prepare_executed = true;
    poly_commit::slice_count = (1 << log_slice_number) + 1;
    poly_commit::slice_size = 1 << (bl + rs_code_rate - log_slice_number);
    poly_commit::slice_real_ele_cnt = poly_commit::slice_size >> rs_code_rate;
    poly_commit::l_eval_len = poly_commit::slice_count * poly_commit::slice_size;
    poly_commit::l_eval = new fieldElement[poly_commit::l_eval_len];
    poly_commit::mask_position_gap = poly_commit::slice_size;          // one mask element: the largest power of two <= slice_size / 1
    poly_prover.all_pri_mask = mask;                                   // mask_size_after_padding == 1
    poly_commit::all_pri_msk_arr = new fieldElement[1];
    poly_commit::all_pri_msk_arr[0] = mask[0];
    init_scratch_pad(poly_commit::slice_size);                         // poly_commit.h:74: FFT scratch of the opening phase
    lap(1);
    const int slice_size = poly_commit::slice_size, half = slice_size / 2;
    // ---- fri.cpp:36-139 (request_init_commit, oracle 0): bookkeeping + arrays
    const int lw = bl + rs_code_rate - log_slice_number;               // log_current_witness_size_per_slice
    fri::__fri_timer = 0;
    fri::current_step_no = 0;
    fri::log_current_witness_size_per_slice = lw;
    fri::witness_bit_length_per_slice = bl - log_slice_number;
    fri::L_group = new fieldElement[1 << lw];
    {
        const fieldElement rou = fieldElement::getRootOfUnity(lw);
        fri::L_group[0] = fieldElement(1);
        for (int i = 1; i < (1 << lw); ++i) fri::L_group[i] = fri::L_group[i - 1] * rou;
    }
    fri::leaf_hash[0] = new __hhash_digest[half];
    fri::witness_merkle[0] = (__hhash_digest *)malloc((size_t)half * 2 * sizeof(__hhash_digest));   // merkle_tree.cpp:17
    merkle_tree::size_after_padding = half;
    ck(vp_commit_export(ctx, mf(poly_commit::l_eval), reinterpret_cast<uint8_t *>(fri::leaf_hash[0]),
                        reinterpret_cast<uint8_t *>(fri::witness_merkle[0])),
       "vp_commit_export");
    lap(2);
    fri::witness_rs_codeword_interleaved[0] = new fieldElement[1 << (bl + rs_code_rate)];
    const int log_leaf_size = log_slice_number + 1;
    ck(vp_commit_export_interleaved(ctx, 0, mf(fri::witness_rs_codeword_interleaved[0])), "vp_commit_export_interleaved");   // (transposed on the device)
    for (int i = 0; i < slice_number; ++i) {                           // fri.cpp:69-96
        fri::witness_rs_codeword_before_arrange[0][i] = &poly_commit::l_eval[i * slice_size];
        fri::witness_rs_mapping[0][i] = new int[1 << lw];
        for (int j = 0; j < half; ++j) {
            const int at = (j << log_leaf_size) | (i << 1);
            fri::witness_rs_mapping[0][i][j] = at;
            fri::witness_rs_mapping[0][i][j + half] = at;
        }
    }
    witness_merkle_size[0] = half;
    fri::visited_init[0] = new bool[1 << lw]();
    fri::visited_witness[0] = new bool[1 << (bl + rs_code_rate)]();
    lap(3);
    poly_prover.total_time = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    if (getenv("VP_TIMING"))
        fprintf(stderr, "virgo_b200 prover: commit_private %.3f ms (device commit %.3f ms; vp_commit_private call %.3f, scratch pad %.3f, L_group + export %.3f, "
                        "interleaving + flags %.3f)\n",
                poly_prover.total_time * 1e3, (double)vp_last_commit_ms(ctx), lap_ms[0], lap_ms[1], lap_ms[2], lap_ms[3]);
    return root;
}

// prover.cpp:532-540
F prover::inner_prod(const vector<F> &a, const vector<F> &b, u64 l) {
    F out;
    ck(vp_dot_host(ctx, cf(a.data()), cf(b.data()), l, mf(&out)), "vp_dot_host");
    return out;
}

// prover.cpp:542-546 -> commit_public_array (lib/virgo/src/poly_commit.h:126-349). The inner product, the encoding of the
// public array, the 2n-point products / inverse FFT / quotient extension per slice, the virtual oracle and the second
// Merkle commitment run on the device (vp_inner_prod, vp_commit_public); as in commit_private, the process-global arrays
// the reference's FRI rounds and verifier read afterwards are allocated and indexed exactly as there and filled from the
// device results (poly_commit.h:131-136,202,303-330; fri.cpp:36-139 for oracle 1).
virgo::__hhash_digest prover::commit_public(vector<F> &pub, F &inner_product_sum, std::vector<F> &mask,
                                            vector<F> &all_sum) {
    using namespace virgo;
    ck(vp_inner_prod(ctx, cf(pub.data()), C.circuit[0].size, mf(&inner_product_sum)), "vp_inner_prod");
    const int bl = C.circuit[0].bitLength;
    bool zero_mask = true;
    for (const F &m : mask) zero_mask = zero_mask && m == F_ZERO;
    if (!gpu_commit || !zero_mask)
        return poly_prover.commit_public_array(mask, pub.data(), bl, inner_product_sum, all_sum.data());
    const auto t0 = std::chrono::high_resolution_clock::now();
    __hhash_digest root_h;
    std::vector<vp_F> sums(slice_number + 1);
    ck(vp_commit_public(ctx, cf(pub.data()), pub.size(), cf(mask.data()), mask.size(), reinterpret_cast<uint8_t *>(&root_h), sums.data()),
       "vp_commit_public");
    const int slice_size = poly_commit::slice_size, slice_count = poly_commit::slice_count, half = slice_size / 2;
    for (int i = 0; i < slice_count; ++i) all_sum[i] = F((long long)sums[i].re, (long long)sums[i].im);
    // ---- poly_commit.h:131-136,202: arrays of the virtual oracle and of h
    fri::virtual_oracle_witness = new fieldElement[slice_size * slice_count];
    fri::virtual_oracle_witness_msk = new fieldElement[slice_size]();
    fri::virtual_oracle_witness_msk_mapping = new int[slice_size];
    fri::virtual_oracle_witness_mapping = new int[slice_size * slice_count];
    poly_commit::q_eval_len = poly_commit::l_eval_len;
    poly_commit::q_eval = new fieldElement[1];                       // (only read inside commit_public_array itself)
    while (mask.size() < (size_t)(slice_size / poly_commit::mask_position_gap)) mask.push_back(F_ZERO);   // :139-140
    poly_commit::all_pub_msk_arr = new fieldElement[mask.size()]();
    poly_commit::h_eval_arr = new fieldElement[slice_count * slice_size];
    fri::leaf_hash[1] = new __hhash_digest[half];
    fri::witness_merkle[1] = (__hhash_digest *)malloc((size_t)half * 2 * sizeof(__hhash_digest));
    ck(vp_commit_public_export(ctx, mf(poly_commit::h_eval_arr), mf(fri::virtual_oracle_witness), reinterpret_cast<uint8_t *>(fri::leaf_hash[1]),
                               reinterpret_cast<uint8_t *>(fri::witness_merkle[1])),
       "vp_commit_public_export");
    const int log_leaf_size = log_slice_number + 1;
    for (int j = 0; j < slice_size; ++j)                              // :258-266 (mask slice: values zero)
        fri::virtual_oracle_witness_msk_mapping[j] = (j < half ? j : j - half) << 1;
    for (int i = 0; i < slice_number; ++i)                            // :310-321
        for (int j = 0; j < slice_size; ++j) {
            const int jj = j < half ? j : j - half;
            fri::virtual_oracle_witness_mapping[jj << log_slice_number | i] = jj << log_leaf_size | (i << 1) | 0;
        }
    // ---- fri.cpp:36-139 (request_init_commit, oracle 1)
    const int lw = bl + rs_code_rate - log_slice_number;
    fri::__fri_timer = 0;
    fri::current_step_no = 0;
    fri::log_current_witness_size_per_slice = lw;
    fri::witness_bit_length_per_slice = bl - log_slice_number;
    merkle_tree::size_after_padding = half;
    fri::witness_rs_codeword_interleaved[1] = new fieldElement[1 << (bl + rs_code_rate)];
    ck(vp_commit_export_interleaved(ctx, 1, mf(fri::witness_rs_codeword_interleaved[1])), "vp_commit_export_interleaved");
    for (int i = 0; i < slice_number; ++i) {
        fri::witness_rs_codeword_before_arrange[1][i] = &poly_commit::h_eval_arr[i * slice_size];
        fri::witness_rs_mapping[1][i] = new int[1 << lw];
        for (int j = 0; j < half; ++j) {
            const int at = (j << log_leaf_size) | (i << 1);
            fri::witness_rs_mapping[1][i][j] = at;
            fri::witness_rs_mapping[1][i][j + half] = at;
        }
    }
    witness_merkle_size[1] = half;
    fri::visited_init[1] = new bool[1 << lw]();
    fri::visited_witness[1] = new bool[1 << (bl + rs_code_rate)]();
    g_fri_ctx = ctx;
    g_device = getenv("VP_DEVICE") ? atoi(getenv("VP_DEVICE")) : 0;   // (the device commit is single-GPU: same choice as in the constructor)
    const double dt = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    poly_prover.total_time += dt;
    if (getenv("VP_TIMING"))
        fprintf(stderr, "virgo_b200 prover: commit_public %.3f ms (device %.3f ms, the rest: copies + the reference's bookkeeping arrays)\n",
                dt * 1e3, (double)vp_last_commit_ms(ctx));
    return root_h;
}
#ifdef VP_DROPIN_FRI
// fri::commit_phase_step (lib/virgo/src/fri.cpp:289-418), called once per fold challenge by
// poly_commit_prover::commit_phase (vpd_verifier.cpp:43-73). The fold of the 64 codewords, the leaf chains and the level's
// Merkle tree run on the device (vp_fri_commit_steps) on the virtual oracle commit_public left there; the arrays the
// reference's query phase reads (fri::request_step_commit, fri.cpp:232-287: cpd.rs_codeword / _msk / their mappings,
// cpd.merkle, visited) are allocated and indexed exactly as there and filled from the device results.
// When the commitments were made by the reference's CPU code (VP_CPU_COMMIT, non-zero masks) so is this step.
virgo::__hhash_digest virgo::fri::commit_phase_step(virgo::fieldElement r) {
    using namespace virgo;
    if (!g_fri_ctx) return commit_phase_step_reference(r);
    const auto t0 = std::chrono::high_resolution_clock::now();
    const int lvl = current_step_no, nxt = (1 << log_current_witness_size_per_slice) / 2, half = nxt / 2;
    const int slice_count = poly_commit::slice_count, log_leaf_size = log_slice_number + 1;
    __hhash_digest root;
    ck(vp_fri_commit_steps(g_fri_ctx, cf(&r), 1, reinterpret_cast<uint8_t *>(&root)), "vp_fri_commit_steps");
    if (cpd.rs_codeword[lvl] == NULL) cpd.rs_codeword[lvl] = new fieldElement[(size_t)nxt * slice_count];          // :290-293
    if (cpd.rs_codeword_msk[lvl] == NULL) cpd.rs_codeword_msk[lvl] = new fieldElement[nxt];
    for (int i = 0; i < nxt; ++i) cpd.rs_codeword_msk[lvl][i] = fieldElement(0);                                  // zero mask: :359-386
    if (cpd.merkle[lvl] == NULL) cpd.merkle[lvl] = (__hhash_digest *)malloc((size_t)half * 2 * sizeof(__hhash_digest));   // merkle_tree.cpp:17
    ck(vp_fri_export_level(g_fri_ctx, lvl, mf(cpd.rs_codeword[lvl]), reinterpret_cast<uint8_t *>(cpd.merkle[lvl])), "vp_fri_export_level");
    for (int i = 0; i < nxt; ++i) L_group[i] = L_group[i * 2];                                                   // :338-340
    cpd.rs_codeword_mapping[lvl] = new int[(size_t)nxt * slice_count];                                            // :343-357
    for (int i = 0; i < half; ++i)
        for (int j = 0; j < slice_number; ++j) {
            const int at = (i << log_leaf_size) | (j << 1);
            cpd.rs_codeword_mapping[lvl][i << log_slice_number | j] = at;
            cpd.rs_codeword_mapping[lvl][(i + half) << log_slice_number | j] = at;
        }
    cpd.rs_codeword_msk_mapping[lvl] = new int[nxt];                                                              // :374-381
    for (int i = 0; i < half; ++i) cpd.rs_codeword_msk_mapping[lvl][i] = cpd.rs_codeword_msk_mapping[lvl][i + half] = i << 1;
    visited[lvl] = new bool[(size_t)nxt * 4 * slice_count]();                                                      // :386-387
    merkle_tree::size_after_padding = half;
    cpd.merkle_size[lvl] = half;
    __fri_timer += std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
    log_current_witness_size_per_slice--;
    current_step_no++;
    return root;
}

// fft_circuit_gkr::fft_gkr (lib/virgo/src/fft_circuit_GKR.cpp:833-849), called once per opening by
// poly_commit_verifier::verify_poly_commitment (vpd_verifier.cpp:92). The reference draws every random value of this inner
// GKR with fieldElement::random() and none depends on a prover message: they are drawn here, in the reference's order and
// number (so the verifier's later draws see the same stream), and the whole protocol runs in vp_fft_gkr.
int virgo::fft_circuit_gkr::fft_gkr(int lg_size, double &vt, int &ps, double &pt) {
    using namespace virgo;
    if (!g_fri_ctx || lg_size < 1 || lg_size > 24) return fft_gkr_reference(lg_size, vt, ps, pt);
    const size_t n_rnd = vp_fft_gkr_rnd_count(lg_size);
    std::vector<fieldElement> rnd(n_rnd);
    for (auto &x : rnd) x = fieldElement::random();
    int ok = 0;
    ps = 0;
    ck(vp_fft_gkr(g_device, lg_size, cf(rnd.data()), n_rnd, nullptr, nullptr, 0, nullptr, &ps, &ok, &vt, &pt, nullptr), "vp_fft_gkr");
    if (!ok) fprintf(stderr, "Error, fft gkr failed\n");   // :843-844
    if (getenv("VP_TIMING")) fprintf(stderr, "virgo_b200 prover: fft_gkr(%d) %.3f ms\n", lg_size, pt * 1e3);
    return 0;
}
#endif
#endif
