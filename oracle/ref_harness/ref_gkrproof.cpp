// Test infrastructure: the reference's (dead) proof container `GKRProof` (/root/reference/src/GKRProof.hpp:10-58,
// 101-140, included from where it lies) reading and re-writing a byte stream produced by the product's
// vp_transcript_to_gkrproof (virgo-plus_b200/host/proof_io.cpp). The header does not compile inside the reference (it names
// NetIO and virgo::poly_commit::PolyProof, which do not exist in its tree), so both get empty stand-ins here; the vectors,
// F, quadratic_poly and the read / write code are the reference's own.
//   usage: ref_gkrproof <file>   -> prints the members GKRProof::read found ("member index count" + "re im" lines),
//                                   the number of bytes it consumed, and whether GKRProof::write reproduces them.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "config_pc.hpp"    // F = virgo::fieldElement
#include "polynomial.h"     // quadratic_poly {a, b, c}

using std::ostream;
using std::vector;
struct NetIO {
    void send_data(const char *, size_t) {}
    void recv_data(char *, size_t) {}
};
namespace virgo { namespace poly_commit {
struct PolyProof {
    void write(ostream &) const {}
    void read(std::istream &) {}
    void send(NetIO *) {}
    void recv(NetIO *) {}
};
} }
#include "GKRProof.hpp"

static void pf(const F &x) { printf("%llu %llu\n", (unsigned long long)x.real, (unsigned long long)x.img); }

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    virgo::fieldElement::init();
    std::ifstream f(argv[1], std::ios::binary);
    std::string bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::istringstream in(bytes);
    GKRProof pr;
    pr.read(in);
    const size_t used = (size_t)in.tellg();
    printf("final_claims_u %zu\n", pr.final_claims_u.size());
    for (auto &x : pr.final_claims_u) pf(x);
    printf("final_claims %zu\n", pr.final_claims.size());
    for (auto &x : pr.final_claims) pf(x);
    printf("final_claims_v %zu\n", pr.final_claims_v.size());
    for (auto &v : pr.final_claims_v) { printf("  row %zu\n", v.size()); for (auto &x : v) pf(x); }
    const vector<vector<quadratic_poly>> *ps[3] = {&pr.polys_u, &pr.polys_v, &pr.polys};
    const char *names[3] = {"polys_u", "polys_v", "polys"};
    for (int k = 0; k < 3; ++k) {
        printf("%s %zu\n", names[k], ps[k]->size());
        for (auto &v : *ps[k]) { printf("  row %zu\n", v.size()); for (auto &q : v) { pf(q.a); pf(q.b); pf(q.c); } }
    }
    std::ostringstream out;
    pr.write(out);
    const std::string w = out.str();
    printf("consumed %zu of %zu bytes; write() reproduces them: %s\n", used, bytes.size(),
           (w.size() == used && w == bytes.substr(0, used)) ? "yes" : "NO");
    return 0;
}
