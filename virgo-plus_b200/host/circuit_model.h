// Host-side circuit model for the B200 GKR prover: layered "unlayered-GKR" circuit with
// per-(layer, source-layer) dad subsets, plus an instance count K for data-parallel replication.
//
// Mirrors (not copies) the reference data model:
//   gate / layer / layeredCircuit      /root/reference/src/circuit.h:11-46
//   gateType numeric values            /root/reference/src/inputCircuit.hpp:14-16
//   subsetInit                         /root/reference/src/circuit.cpp:43-80
//   .pws parser + DAG_to_layered       /root/reference/src/main.cpp:15-137,161-231
//
// Storage is struct-of-arrays with 32-bit in-layer indices (the GPU form); a circuit with
// `instances = K > 1` is the template of one instance and stands for the K-fold instance-major
// replication described in SURVEY.md 9.3:  g = k*S_i + g0, u = k*S_{i-1} + u0, v = k*S_l + v0,
// lv = (K-1-k)*D_l + lv0, dadId[(K-1-k)*D_l + j] = k*S_l + dadId0[j].
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../csrc/field.cuh"

namespace vp {

enum GateType : uint8_t {
    Mul = 0, Add = 1, Sub = 2, AntiSub = 3, Naab = 4, AntiNaab = 5, Input = 6,
    Mulc = 7, Addc = 8, Xor = 9, Not = 10, Copy = 11, NUM_GATE_TYPES = 12
};

inline bool is_binary(uint8_t ty) {
    return ty == Mul || ty == Add || ty == Sub || ty == AntiSub || ty == Naab || ty == AntiNaab || ty == Xor;
}

struct Layer {
    uint64_t size = 0;   // gates in ONE instance
    // gate fields, one entry per gate of one instance (layer 0: ty == Input, u unused)
    std::vector<uint8_t> ty;
    std::vector<int32_t> l;       // source layer of v; -1 for unary gates
    std::vector<uint32_t> u, v, lv;
    std::vector<F> c;             // constant of Addc / Mulc (empty if the layer has none)
    std::vector<uint8_t> is_assert;  // empty if the layer has none
    // dad subsets (per source layer l < i), one instance
    std::vector<std::vector<uint32_t>> dadId;
    std::vector<uint64_t> dadSize;
};

int ceil_log2(uint64_t x);  // reference rule: (int)log2(x) with fix-up; x == 0 -> -1 here

struct Circuit {
    std::vector<Layer> layers;
    uint64_t instances = 1;
    std::vector<uint64_t> inputs;  // instances * layers[0].size values (< p), instance-major

    int n_layers() const { return (int)layers.size(); }
    uint64_t layer_size(int i) const { return layers[i].size * instances; }          // replicated size
    int bit_length(int i) const { return ceil_log2(layer_size(i)); }
    uint64_t dad_size(int i, int l) const { return layers[i].dadSize[l] * instances; }
    // -1 for an empty subset (the reference stores INT_MIN there, circuit.cpp:73)
    int dad_bit_length(int i, int l) const { return ceil_log2(dad_size(i, l)); }
    int max_dad_bit_length(int i) const;
    int max_bit_length() const;
    uint64_t total_gates() const;  // non-input gates, all instances

    // Build dad subsets + lv for the template (instances must be 1 when called).
    void subset_init();
    // Materialise the K instances into a flat circuit (instances = 1), re-deriving the subsets from
    // scratch with subset_init(); used to cross-check the replication index rules.
    Circuit expand() const;
    // Same template, K instances; inputs are drawn like the reference draws them at parse time.
    Circuit replicate(uint64_t K) const;
    void draw_inputs_like_reference();  // glibc random() % p, default seed, file order (main.cpp:188)
    std::string validate() const;       // "" if ok, else a description of the first problem
};

// .pws loader (grammar: SURVEY.md 9.6). Reproduces the reference's layering and its quirks.
// Returns "" on success, else an error string.
std::string load_pws(const std::string& path, Circuit& out);
std::string load_pws_text(const char* text, size_t len, Circuit& out);

// Synthetic unlayered circuit with the semantics of layeredCircuit::randomize (circuit.cpp:17-41),
// from an own seeded SplitMix64 stream.
Circuit random_circuit(int n_layers, int log_size, uint64_t seed);

// ------------------------------------------------------------------ challenge stream
// The verifier's challenges in the exact order verifier.cpp:134-337 draws them, from glibc
// random() after srand(seed) (fieldElement.cpp:106-124,362-367), without touching the process
// global RNG state.
struct GlibcRandom {
    GlibcRandom(unsigned seed);
    long next();      // == random()
    uint64_t number();  // == fieldElement::randomNumber()
    F field();          // == fieldElement::random()
    char state[128];
    char buf[64];  // struct random_data storage (opaque; sized generously)
};

struct LayerChallenges {
    std::vector<F> r_u;   // max_bl entries (only bl(i-1) used)
    F assert_random;
    std::vector<F> r_v;   // maxDadBitLength(i) entries (empty if phase 2 is skipped)
    std::vector<F> sig;   // n_layers entries
    std::vector<F> r_liu; // max_bl entries
};
struct ChallengeStream {
    std::vector<F> r_out;                  // bl(out) entries
    std::vector<LayerChallenges> layer;    // indexed by layer id (1..n-1); [0] unused
    uint64_t count = 0;
};
ChallengeStream draw_challenges(const Circuit& c, unsigned seed);

}  // namespace vp
