// TEST INFRASTRUCTURE ONLY (oracle/_ref build): runs the UNMODIFIED reference poly_commit_prover::commit_private_array
// (lib/virgo/src/poly_commit.h:41-124 -> vpd_prover.cpp:9-14 -> fri.cpp:36-139 -> merkle_tree.cpp:7-51, SHA3 from the
// reference's prebuilt libXKCP.a) on an array read from a file and dumps what it produced:
//   usage: ref_pc_commit <log_len> <array.bin> <out.bin> [<public.bin> <out2.bin> [<randomness.bin> <out3.bin>]]
//   array.bin: 2^log_len field elements {u64 real, u64 img};  the mask is the GKR prover's: one zero (prover.cpp:524-530)
//   out.bin  : root[32] | l_eval[65 * slice_size * 16] | leaf_hash[slice_size/2 * 32] | merkle tree[slice_size * 32]
// and prints the commit time the reference accounts for itself (poly_prover.total_time).
// With public.bin (2^log_len elements: the verifier's eq table in the real program, verifier.cpp:367-381) it goes on with
// commit_public_array (poly_commit.h:126-349; public mask = one zero, target sum = <array, public>) and dumps
//   out2.bin : root_h[32] | all_sum[65 * 16] | h_eval_arr[65 * slice_size * 16] | virtual_oracle_witness[64 * slice_size * 16]
//              | virtual_oracle_witness_msk[slice_size * 16]
// With randomness.bin (log_len - 6 field elements) it runs the FRI commit phase as poly_commit_prover::commit_phase does
// (vpd_verifier.cpp:43-73), with the given fold challenges in place of fieldElement::random(): one fri::commit_phase_step
// (fri.cpp:289-418) per element, and dumps
//   out3.bin : roots[steps * 32] | per level l: rs_codeword[64 * (slice_size >> (l+1)) * 16] | merkle[(slice_size >> (l+1)) * 32]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "virgo/src/fri.h"
#include "virgo/src/poly_commit.h"

using namespace virgo;

int main(int argc, char **argv) {
    if (argc < 4) return 2;
    const int log_len = atoi(argv[1]);
    const size_t n = (size_t)1 << log_len;
    std::vector<fieldElement> arr(n);
    FILE *f = fopen(argv[2], "rb");
    if (!f || fread(arr.data(), sizeof(fieldElement), n, f) != n) return 3;
    fclose(f);
    fieldElement::init();
    poly_commit::poly_commit_prover p;
    std::vector<fieldElement> mask(1, fieldElement::zero());
    __hhash_digest root = p.commit_private_array(arr.data(), log_len, mask);
    const size_t slice_size = (size_t)poly_commit::slice_size, half = slice_size / 2;
    f = fopen(argv[3], "wb");
    if (!f) return 4;
    fwrite(&root, 32, 1, f);
    fwrite(poly_commit::l_eval, sizeof(fieldElement), (size_t)poly_commit::slice_count * slice_size, f);
    fwrite(fri::leaf_hash[0], 32, half, f);
    fwrite(fri::witness_merkle[0], 32, slice_size, f);
    fclose(f);
    printf("slice_size %zu commit_seconds %.6f\n", slice_size, p.total_time);
    if (argc >= 6) {
        std::vector<fieldElement> pub(n);
        f = fopen(argv[4], "rb");
        if (!f || fread(pub.data(), sizeof(fieldElement), n, f) != n) return 5;
        fclose(f);
        fieldElement target = fieldElement::zero();
        for (size_t i = 0; i < n; ++i) target = target + arr[i] * pub[i];
        std::vector<fieldElement> pub_mask(1, fieldElement::zero()), all_sum(poly_commit::slice_count);
        const double t_before = p.total_time;
        __hhash_digest root_h = p.commit_public_array(pub_mask, pub.data(), log_len, target, all_sum.data());
        f = fopen(argv[5], "wb");
        if (!f) return 6;
        fwrite(&root_h, 32, 1, f);
        fwrite(all_sum.data(), sizeof(fieldElement), all_sum.size(), f);
        fwrite(poly_commit::h_eval_arr, sizeof(fieldElement), (size_t)poly_commit::slice_count * slice_size, f);
        fwrite(fri::virtual_oracle_witness, sizeof(fieldElement), (size_t)(poly_commit::slice_count - 1) * slice_size, f);
        fwrite(fri::virtual_oracle_witness_msk, sizeof(fieldElement), slice_size, f);
        fclose(f);
        printf("commit_public_seconds %.6f\n", p.total_time - t_before);
        if (argc >= 8) {
            const int steps = log_len - log_slice_number;
            std::vector<fieldElement> rnd(steps);
            f = fopen(argv[6], "rb");
            if (!f || fread(rnd.data(), sizeof(fieldElement), steps, f) != (size_t)steps) return 7;
            fclose(f);
            std::vector<__hhash_digest> roots(steps);
            const double t_init = fri::__fri_timer;                  // request_init_commit(.., 1) left its own time there
            for (int s = 0; s < steps; ++s) roots[s] = fri::commit_phase_step(rnd[s]);
            f = fopen(argv[7], "wb");
            if (!f) return 8;
            fwrite(roots.data(), 32, steps, f);
            for (int s = 0; s < steps; ++s) {
                const size_t m = slice_size >> (s + 1);
                fwrite(fri::cpd.rs_codeword[s], sizeof(fieldElement), 64 * m, f);
                __hhash_digest zero;
                memset(&zero, 0, sizeof zero);
                fwrite(&zero, 32, 1, f);                             // node 0 of the heap is never read
                fwrite(fri::cpd.merkle[s] + 1, 32, m - 1, f);
            }
            fclose(f);
            printf("fri_commit_seconds %.6f\n", fri::__fri_timer - t_init);
        }
    }
    return 0;
}
