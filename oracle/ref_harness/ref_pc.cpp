// TEST INFRASTRUCTURE ONLY (oracle/_ref build): runs the UNMODIFIED reference poly_commit_prover::commit_private_array
// (lib/virgo/src/poly_commit.h:41-124 -> vpd_prover.cpp:9-14 -> fri.cpp:36-139 -> merkle_tree.cpp:7-51, SHA3 from the
// reference's prebuilt libXKCP.a) on an array read from a file and dumps what it produced:
//   usage: ref_pc_commit <log_len> <array.bin> <out.bin>
//   array.bin: 2^log_len field elements {u64 real, u64 img};  the mask is the GKR prover's: one zero (prover.cpp:524-530)
//   out.bin  : root[32] | l_eval[65 * slice_size * 16] | leaf_hash[slice_size/2 * 32] | merkle tree[slice_size * 32]
// and prints the commit time the reference accounts for itself (poly_prover.total_time).
#include <cstdio>
#include <cstdlib>
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
    return 0;
}
