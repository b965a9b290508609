// Transcript containers (host side).
//  * GKRProof byte stream: the layout of GKRProof::write in the reference's (dead) src/GKRProof.hpp:23-58 --
//    for each member in the order final_claims_u, final_claims, final_claims_v, polys_u, polys_v, polys:
//    u64 count, then (nested vectors: per inner vector u64 len, then) the raw elements; F = 16 B {real, img}
//    little endian, quadratic_poly = 48 B {a, b, c}. Outer vectors are indexed by layer id (n_layers entries,
//    entry 0 empty). The reference's poly_proof part (a type that does not exist in its tree) is replaced by
//    a two-element trailer: u64 2, Vres, input-layer MLE.
//  * text dump: "TAG real img" per line in emission order: VRES; per round CH,PA,PB,PC;
//    CH,CLAIM_U; CLAIM_V x layer; CH,CLAIM_LIU; INPUT_MLE.
#pragma once
#include <string>
#include <vector>

#include "circuit_model.h"

namespace vp {
std::vector<unsigned char> transcript_to_gkrproof(const Circuit& c, const F* transcript);
// returns "" on success; transcript must have vp_transcript_len entries
std::string gkrproof_to_transcript(const Circuit& c, const unsigned char* bytes, size_t len, F* transcript);
std::string transcript_text(const Circuit& c, const F* transcript, const F* challenges);
size_t transcript_len(const Circuit& c);
}  // namespace vp
