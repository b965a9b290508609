// TEST INFRASTRUCTURE ONLY (oracle/_ref build): drives the UNMODIFIED reference fft_circuit_GKR (lib/virgo/src/
// fft_circuit_GKR.cpp, compiled into this translation unit from where it lies so that its internal functions and globals
// are reachable) in the order of its own engage_gkr (:784-831) and records what can be observed from outside:
//   usage: ref_fftgkr <lg> <seed> <out.bin>
//   out.bin: n_rnd (u64) | rnd[n_rnd] (the fieldElement::random() stream the run consumed, regenerated from the same seed)
//            | layers: E, F_{lg-1}..F_0, S (2^lg each), P (64 * 2^lg), O (64)
//            | claims: a_0, after addition_layer, after mult_layer, after intermediate_layer, after ifft_gkr, alpha, beta
//            | proof_size (u64) | ok (u64) | then fft_gkr's own {ps} from a second, plain call with the same seed (u64)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "virgo/src/fft_circuit_GKR.cpp"

using namespace virgo;
using namespace virgo::fft_circuit_gkr;

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    const int lg = atoi(argv[1]);
    const unsigned seed = (unsigned)atoi(argv[2]);
    fieldElement::init();
    const size_t n = (size_t)1 << lg;
    const size_t n_rnd = (size_t)lg + 64 + 2 * (lg + 10) + 2 * (lg + 6) + 2 * lg + (size_t)lg * (2 * lg + 2);
    srand(seed);
    std::vector<fieldElement> rnd(n_rnd);
    for (auto &x : rnd) x = fieldElement::random();
    srand(seed);
    // ---- fft_gkr :833-849 up to engage_gkr
    v_time = 0; proof_size = 0; p_time = 0;
    init_array(lg + 6);
    fieldElement *r = new fieldElement[lg];
    for (int i = 0; i < lg; ++i) r[i] = fieldElement::random();
    build_circuit(lg, r);
    // ---- engage_gkr :784-831, with the running claim recorded between the stages
    alpha = fieldElement(1);
    beta = fieldElement(0);
    fieldElement *r_0 = new fieldElement[lg + 10], *om_r_0 = new fieldElement[lg + 10], *r_1 = new fieldElement[lg + 10], *om_r_1 = new fieldElement[lg + 10];
    fieldElement *r_u = new fieldElement[lg + 10], *om_r_u = new fieldElement[lg + 10], *r_v = new fieldElement[lg + 10], *om_r_v = new fieldElement[lg + 10];
    refresh_randomness(r_0, om_r_0, lg + 10);
    refresh_randomness(r_1, om_r_1, lg + 10);
    const size_t L = C.circuit_val.size();
    std::vector<fieldElement> claims;
    fieldElement abs_ = V_output(om_r_0, r_0, C.circuit_val[L - 1], mylog(C.size[L - 1]), C.size[L - 1]);
    claims.push_back(abs_);
    bool ok = true;
    ok &= addition_layer(abs_, C.circuit_val[L - 2], C.size[L - 2] / 64, 64, r_0, om_r_0, r_1, om_r_1, r_u, om_r_u, r_v, om_r_v);
    claims.push_back(abs_);
    ok &= mult_layer(abs_, C.circuit_val[L - 3], C.size[L - 3], 64, r_0, om_r_0, r_1, om_r_1, r_u, om_r_u, r_v, om_r_v);
    claims.push_back(abs_);
    ok &= intermediate_layer(abs_, C.circuit_val[L - 4], C.size[L - 4], r_0, om_r_0, r_1, om_r_1, r_u, om_r_u, r_v, om_r_v);
    claims.push_back(abs_);
    ok &= ifft_gkr(abs_, C.size[L - 4], r_0, om_r_0, r_1, om_r_1, r_u, om_r_u, r_v, om_r_v);
    claims.push_back(abs_);
    ok &= extension_gkr(abs_, C.size[L - 4], r_0, om_r_0, r_1, om_r_1, r_u, om_r_u, r_v, om_r_v);
    claims.push_back(alpha);
    claims.push_back(beta);
    FILE *f = fopen(argv[3], "wb");
    if (!f) return 3;
    unsigned long long w = n_rnd;
    fwrite(&w, 8, 1, f);
    fwrite(rnd.data(), sizeof(fieldElement), n_rnd, f);
    for (int t = lg; t <= 2 * lg + 1; ++t) fwrite(C.circuit_val[t], sizeof(fieldElement), n, f);   // E, the butterfly layers, S
    fwrite(C.circuit_val[2 * lg + 2], sizeof(fieldElement), 64 * n, f);
    fwrite(C.circuit_val[2 * lg + 3], sizeof(fieldElement), 64, f);
    fwrite(claims.data(), sizeof(fieldElement), claims.size(), f);
    w = (unsigned long long)proof_size; fwrite(&w, 8, 1, f);
    w = ok ? 1 : 0; fwrite(&w, 8, 1, f);
    // the stock entry point on the same stream: its proof size must be the one counted above
    C.circuit_val.clear();
    C.size.clear();
    srand(seed);
    double vt = 0, pt = 0;
    int ps = 0;
    fft_gkr(lg, vt, ps, pt);
    w = (unsigned long long)ps; fwrite(&w, 8, 1, f);
    fclose(f);
    printf("lg %d ok %d proof_size %d fft_gkr: ps %d prover_seconds %.6f verifier_seconds %.6f\n", lg, (int)ok, proof_size, ps, pt, vt);
    return 0;
}
