// Host circuit model: .pws loader, layering, dad subsets, replication, challenge stream.
// See circuit_model.h for the reference files each piece mirrors.
#include "circuit_model.h"

#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <deque>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace vp {

int ceil_log2(uint64_t x) {
    // main.cpp:133-136 / circuit.cpp:73-75: (int)log2(size), +1 if 2^that < size.
    if (x == 0) return -1;  // reference: INT_MIN (UB of (int)log2(0)); callers treat both as "empty"
    int b = 63 - __builtin_clzll(x);
    if ((1ULL << b) < x) ++b;
    return b;
}

int Circuit::max_dad_bit_length(int i) const {
    int m = -1;
    for (int l = 0; l < i; ++l) m = std::max(m, dad_bit_length(i, l));
    return m;
}

int Circuit::max_bit_length() const {
    int m = 0;
    for (int i = 0; i < n_layers(); ++i) m = std::max(m, bit_length(i));
    return m;
}

uint64_t Circuit::total_gates() const {
    uint64_t t = 0;
    for (int i = 1; i < n_layers(); ++i) t += layer_size(i);
    return t;
}

// circuit.cpp:43-80. Layers top-down, gates in DESCENDING index order; the first visit of (l, v)
// in that order takes the next subset slot. Unary gates (l == -1) are skipped and keep lv = 0.
void Circuit::subset_init() {
    const int n = n_layers();
    std::vector<std::vector<int>> visited(n);
    std::vector<std::vector<uint32_t>> slot(n);
    for (int i = 0; i < n; ++i) {
        visited[i].assign(layers[i].size, 0);
        slot[i].assign(layers[i].size, 0);
        layers[i].dadId.assign(i, {});
        layers[i].dadSize.assign(i, 0);
        layers[i].lv.assign(layers[i].size, 0);
    }
    for (int i = n - 1; i > 0; --i) {
        Layer& L = layers[i];
        for (uint64_t j = L.size; j-- > 0;) {
            int l = L.l[j];
            if (l < 0) continue;
            uint32_t v = L.v[j];
            if (visited[l][v] != i) {
                visited[l][v] = i;
                slot[l][v] = (uint32_t)L.dadSize[l]++;
                L.dadId[l].push_back(v);
            }
            L.lv[j] = slot[l][v];
        }
    }
}

Circuit Circuit::expand() const {
    Circuit o;
    const uint64_t K = instances;
    o.instances = 1;
    o.layers.resize(layers.size());
    for (size_t i = 0; i < layers.size(); ++i) {
        const Layer& T = layers[i];
        Layer& L = o.layers[i];
        const uint64_t S = T.size;
        L.size = S * K;
        L.ty.resize(L.size);
        L.l.resize(L.size);
        L.u.resize(L.size);
        L.v.resize(L.size);
        if (!T.c.empty()) L.c.resize(L.size);
        if (!T.is_assert.empty()) L.is_assert.resize(L.size);
        for (uint64_t k = 0; k < K; ++k)
            for (uint64_t g = 0; g < S; ++g) {
                uint64_t G = k * S + g;
                L.ty[G] = T.ty[g];
                L.l[G] = T.l[g];
                if (i == 0) {
                    L.u[G] = 0;
                    L.v[G] = 0;
                } else {
                    L.u[G] = (uint32_t)(k * layers[i - 1].size + T.u[g]);
                    L.v[G] = T.l[g] >= 0 ? (uint32_t)(k * layers[T.l[g]].size + T.v[g]) : 0;
                }
                if (!T.c.empty()) L.c[G] = T.c[g];
                if (!T.is_assert.empty()) L.is_assert[G] = T.is_assert[g];
            }
    }
    o.inputs = inputs;
    o.subset_init();
    return o;
}

void Circuit::draw_inputs_like_reference() {
    // main.cpp:188: buildInput(tgt, random() % mod) once per input line, in file order, before
    // srand(3396) -> the process-default generator (TYPE_3, seed 1).
    struct random_data rd;
    char state[128];
    memset(&rd, 0, sizeof rd);
    memset(state, 0, sizeof state);
    initstate_r(1u, state, sizeof state, &rd);
    uint64_t n = instances * layers[0].size;
    inputs.resize(n);
    for (uint64_t i = 0; i < n; ++i) {
        int32_t r;
        random_r(&rd, &r);
        inputs[i] = (uint64_t)r % P;
    }
}

Circuit Circuit::replicate(uint64_t K) const {
    for (const Layer& L : layers)   // before anything is allocated for K instances
        if (K > (1ULL << 31) || L.size * K > (1ULL << 31)) throw std::length_error("replicate: a layer would exceed 2^31 gates");
    Circuit o = *this;
    o.instances = K;
    o.draw_inputs_like_reference();
    return o;
}

std::string Circuit::validate() const {
    char buf[256];
    if (layers.empty()) return "circuit has no layers";
    if (instances == 0) return "instances == 0";
    for (int i = 0; i < n_layers(); ++i) {
        const Layer& L = layers[i];
        if (L.size == 0) {
            snprintf(buf, sizeof buf, "layer %d is empty", i);
            return buf;
        }
        if (instances > (1ULL << 31) || L.size > (1ULL << 31) || L.size * instances > (1ULL << 31)) {
            snprintf(buf, sizeof buf, "layer %d: %llu gates exceed the 2^31 per-layer limit", i,
                     (unsigned long long)(L.size * instances));
            return buf;
        }
        if (L.ty.size() != L.size || L.l.size() != L.size || L.u.size() != L.size || L.v.size() != L.size ||
            L.lv.size() != L.size)
            return "gate arrays have inconsistent lengths";
        if ((int)L.dadId.size() != i || (int)L.dadSize.size() != i) return "dad subsets not initialised";
        for (uint64_t g = 0; g < L.size; ++g) {
            uint8_t ty = L.ty[g];
            if (i == 0) {
                if (ty != Input) return "layer 0 holds a non-input gate";
                continue;
            }
            if (ty >= NUM_GATE_TYPES || ty == Input) {
                snprintf(buf, sizeof buf, "layer %d gate %llu: bad gate type %d", i, (unsigned long long)g, ty);
                return buf;
            }
            if (L.u[g] >= layers[i - 1].size) {
                // e.g. the Not/Copy fall-through (main.cpp:104-110) stored a raw DAG id that is not
                // an in-layer id: the reference would read out of bounds here.
                snprintf(buf, sizeof buf, "layer %d gate %llu: u=%u out of range of layer %d (size %llu)", i,
                         (unsigned long long)g, L.u[g], i - 1, (unsigned long long)layers[i - 1].size);
                return buf;
            }
            if (is_binary(ty)) {
                int l = L.l[g];
                if (l < 0 || l >= i || L.v[g] >= layers[l].size) {
                    snprintf(buf, sizeof buf, "layer %d gate %llu: bad (l,v)=(%d,%u)", i, (unsigned long long)g, l,
                             L.v[g]);
                    return buf;
                }
                if (L.lv[g] >= L.dadSize[l] || L.dadId[l][L.lv[g]] != L.v[g]) {
                    snprintf(buf, sizeof buf, "layer %d gate %llu: lv inconsistent with dadId", i,
                             (unsigned long long)g);
                    return buf;
                }
            } else if (L.l[g] != -1) {
                snprintf(buf, sizeof buf, "layer %d gate %llu: unary gate with l=%d", i, (unsigned long long)g, L.l[g]);
                return buf;
            }
        }
        for (int l = 0; l < i; ++l) {
            if (L.dadId[l].size() != L.dadSize[l]) return "dadId/dadSize mismatch";
            for (uint32_t x : L.dadId[l])   // also entries no gate references: phase 2 / Liu gather through every slot
                if (x >= layers[l].size) {
                    snprintf(buf, sizeof buf, "layer %d: dadId[%d] entry %u out of range of layer %d", i, l, x, l);
                    return buf;
                }
        }
    }
    if (inputs.size() != instances * layers[0].size) return "inputs length != instances * layer-0 size";
    for (uint64_t x : inputs)
        if (x >= P) return "input value >= p";
    return "";
}

// ------------------------------------------------------------------ .pws loader
namespace {

struct DagGate {
    uint8_t ty = 0xff;  // 0xff = hole (id never defined)
    char k0 = 'N', k1 = 'N';  // operand kinds: 'V' variable, 'S' constant, 'N' none
    uint64_t in0 = 0, in1 = 0;
};

// Strict matcher for one line; mirrors the eight std::regex patterns of main.cpp:161-168
// ("P V<t> = V<a> OP V<b> E", "P V<t> = I<k> E", "P O<t> = V<a> E"). Anything else is ignored,
// as in the reference's Release build (assert(false) compiled out, main.cpp:204).
static thread_local uint64_t dag_id_cap = 0;        // number of lines of the file: dense ids cannot exceed it
static thread_local bool dag_overflow = false;      // a number too large to be an id / index was seen
// [0-9]+ like the reference's regexes (any number of digits, leading zeros allowed). A value too large to be an id sets
// `big` and keeps consuming digits: whether that matters is decided once the WHOLE line is known to be a statement (the
// reference skips a malformed line whatever numbers it holds, main.cpp:176-204).
bool read_uint(const char*& p, const char* e, uint64_t& out, bool& big) {
    if (p >= e || *p < '0' || *p > '9') return false;
    uint64_t x = 0;
    while (p < e && *p >= '0' && *p <= '9') {
        if (x > (1ULL << 40)) big = true;   // no silent wrap-around
        else x = x * 10 + (uint64_t)(*p - '0');
        ++p;
    }
    out = x;
    return true;
}
bool eat(const char*& p, const char* e, const char* lit) {
    size_t n = strlen(lit);
    if ((size_t)(e - p) < n || memcmp(p, lit, n) != 0) return false;
    p += n;
    return true;
}

void parse_line(const char* p, const char* e, std::vector<DagGate>& dag, uint64_t& n_inputs_seen,
                std::vector<uint64_t>& input_order) {
    uint64_t tgt, a, b;
    bool big = false, big_ignored = false;
    if (!eat(p, e, "P ")) return;
    if (eat(p, e, "O")) {  // output line: parsed and dropped
        return;
    }
    if (!eat(p, e, "V") || !read_uint(p, e, tgt, big) || !eat(p, e, " = ")) return;
    DagGate g;
    if (eat(p, e, "I")) {
        if (!read_uint(p, e, a, big_ignored) || !eat(p, e, " E") || p != e) return;   // the input index is never used (main.cpp:184-185)
        g.ty = Input;
        g.k0 = 'S';
        g.k1 = 'N';
        ++n_inputs_seen;
        input_order.push_back(tgt);
    } else {
        if (!eat(p, e, "V") || !read_uint(p, e, a, big) || !eat(p, e, " ")) return;
        uint8_t ty;
        if (eat(p, e, "+ ")) ty = Add;
        else if (eat(p, e, "* ")) ty = Mul;
        else if (eat(p, e, "XOR ")) ty = Xor;
        else if (eat(p, e, "NAAB ")) ty = Naab;
        else if (eat(p, e, "minus ")) ty = Sub;
        else if (eat(p, e, "NOT ")) ty = Not;
        else return;
        if (!eat(p, e, "V") || !read_uint(p, e, b, ty == Not ? big_ignored : big) || !eat(p, e, " E") || p != e) return;   // NOT: second operand unused
        g.ty = ty;
        g.k0 = 'V';
        g.in0 = a;
        if (ty == Not) {  // buildGate(Not, tgt, src0, 0, true): second operand is the constant 0
            g.k1 = 'S';
            g.in1 = 0;
        } else {
            g.k1 = 'V';
            g.in1 = b;
        }
    }
    // ids must be dense (dag_to_layered rejects holes), so an id far beyond the lines seen so far can never become
    // valid: refuse it here instead of allocating tgt + 1 entries for a hostile file
    if (big || tgt > dag_id_cap) { dag_overflow = true; return; }   // a well-formed statement with an impossible id: the file is rejected
    if (tgt >= dag.size()) dag.resize(tgt + 1);
    dag[tgt] = g;
}

// main.cpp:15-137 (DAG_to_layered) on the parsed DAG.
std::string dag_to_layered(const std::vector<DagGate>& dag, Circuit& out) {
    const uint64_t n = dag.size();
    char buf[200];
    for (uint64_t i = 0; i < n; ++i) {
        if (dag[i].ty == 0xff) {
            snprintf(buf, sizeof buf, "V%llu is never defined (ids must be dense)", (unsigned long long)i);
            return buf;
        }
        if ((dag[i].k0 == 'V' && dag[i].in0 >= n) || (dag[i].k1 == 'V' && dag[i].in1 >= n)) {
            snprintf(buf, sizeof buf, "V%llu uses an undefined operand", (unsigned long long)i);
            return buf;
        }
    }
    std::vector<uint32_t> in_deg(n, 0);
    std::vector<int> lyr(n, 0);
    // CSR of DAG edges (operand -> user), users in ascending id order like the reference's
    // per-node push_back order.
    std::vector<uint64_t> eoff(n + 1, 0);
    for (uint64_t i = 0; i < n; ++i) {
        if (dag[i].k0 == 'V') ++eoff[dag[i].in0 + 1];
        if (dag[i].k1 == 'V') ++eoff[dag[i].in1 + 1];
    }
    for (uint64_t i = 0; i < n; ++i) eoff[i + 1] += eoff[i];
    std::vector<uint64_t> edges(eoff[n]), fill(eoff.begin(), eoff.end() - 1);
    std::deque<uint64_t> q;
    for (uint64_t i = 0; i < n; ++i) {
        if (dag[i].k0 == 'V') {
            ++in_deg[i];
            edges[fill[dag[i].in0]++] = i;
        }
        if (dag[i].k1 == 'V') {
            ++in_deg[i];
            edges[fill[dag[i].in1]++] = i;
        }
        if (dag[i].ty == Input) q.push_back(i);
    }
    int max_lyr = 0;
    uint64_t visited = 0;
    while (!q.empty()) {
        uint64_t u = q.front();
        q.pop_front();
        ++visited;
        max_lyr = std::max(max_lyr, lyr[u]);
        for (uint64_t k = eoff[u]; k < eoff[u + 1]; ++k) {
            uint64_t v = edges[k];
            if (--in_deg[v] == 0) {
                q.push_back(v);
                lyr[v] = std::max(lyr[v], lyr[u] + 1);
            }
        }
    }
    if (visited != n) return "circuit has a cycle or a gate unreachable from the inputs";

    out = Circuit();
    out.layers.resize(max_lyr + 1);
    std::vector<uint32_t> id_in_lyr(n);
    for (uint64_t i = 0; i < n; ++i) id_in_lyr[i] = (uint32_t)out.layers[lyr[i]].size++;
    for (auto& L : out.layers) {
        L.ty.assign(L.size, 0);
        L.l.assign(L.size, -1);
        L.u.assign(L.size, 0);
        L.v.assign(L.size, 0);
        L.lv.assign(L.size, 0);
    }
    for (uint64_t i = 0; i < n; ++i) {
        const DagGate& g = dag[i];
        const int lg = lyr[i];
        Layer& L = out.layers[lg];
        const uint32_t gid = id_in_lyr[i];
        uint64_t in0 = g.in0, in1 = g.in1;
        uint8_t ty = g.ty;
        switch (g.ty) {
            case Mul: case Add: case Xor: case Sub: case Naab: {
                uint32_t u = id_in_lyr[in0], v = id_in_lyr[in1];
                if (lyr[in0] < lg - 1) {  // make u the operand that lives in layer lg-1
                    std::swap(u, v);
                    std::swap(in0, in1);
                    if (g.ty == Sub) ty = AntiSub;
                    if (g.ty == Naab) ty = AntiNaab;
                }
                L.ty[gid] = ty;
                L.l[gid] = lyr[in1];
                L.u[gid] = u;
                L.v[gid] = v;
                break;
            }
            case Not: case Copy: {
                // main.cpp:104-110: `case Not: case Copy:` has no break and falls into `case Input:`,
                // so the stored u is the RAW DAG id of the operand, not its in-layer id. Reproduced
                // on purpose; Circuit::validate() rejects circuits where that id is out of range.
                if (in0 > 0xffffffffULL) return "NOT operand id exceeds 32 bits";
                L.ty[gid] = ty;
                L.l[gid] = -1;
                L.u[gid] = (uint32_t)in0;
                break;
            }
            case Input:
                L.ty[gid] = Input;
                L.l[gid] = -1;
                break;
            default:
                return "unsupported gate type in .pws";
        }
    }
    return "";
}

}  // namespace

std::string load_pws_text(const char* text, size_t len, Circuit& out) {
    std::vector<DagGate> dag;
    uint64_t n_inputs = 0;
    std::vector<uint64_t> input_order;
    const char* p = text;
    const char* end = text + len;
    dag_overflow = false;
    dag_id_cap = 1;
    for (const char* q = text; q < end; ++q) dag_id_cap += (*q == '\n');
    while (p < end) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
        const char* e = nl ? nl : end;
        parse_line(p, e, dag, n_inputs, input_order);
        p = nl ? nl + 1 : end;
    }
    if (dag_overflow) return "a gate id exceeds the number of lines (ids must be dense) or a number is out of range";
    if (dag.empty()) return "no gates parsed";
    std::string err = dag_to_layered(dag, out);
    if (!err.empty()) return err;
    out.instances = 1;
    out.subset_init();
    // Inputs: one random() % p per input LINE in file order (a redefinition of the same id draws
    // again and the last value wins, as in the reference).
    {
        struct random_data rd;
        char state[128];
        memset(&rd, 0, sizeof rd);
        memset(state, 0, sizeof state);
        initstate_r(1u, state, sizeof state, &rd);
        // in-layer id of input gate = rank among layer-0 gates in DAG-id order
        std::vector<uint64_t> val(dag.size(), 0);
        for (uint64_t tgt : input_order) {
            int32_t r;
            random_r(&rd, &r);
            val[tgt] = (uint64_t)r % P;
        }
        out.inputs.clear();
        for (uint64_t i = 0; i < dag.size(); ++i)
            if (dag[i].ty == Input) out.inputs.push_back(val[i]);
    }
    return out.validate();
}

std::string load_pws(const std::string& path, Circuit& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return "cannot open " + path;
    std::stringstream ss;
    ss << f.rdbuf();
    std::string s = ss.str();
    return load_pws_text(s.data(), s.size(), out);
}

// ------------------------------------------------------------------ synthetic circuit
namespace {
struct SplitMix64 {
    uint64_t s;
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
};
}  // namespace

Circuit random_circuit(int n_layers, int log_size, uint64_t seed) {
    Circuit c;
    SplitMix64 rng{seed};
    const uint64_t S = 1ULL << log_size;
    c.layers.resize(n_layers);
    for (int i = 0; i < n_layers; ++i) {
        Layer& L = c.layers[i];
        L.size = S;
        L.ty.assign(S, Input);
        L.l.assign(S, -1);
        L.u.assign(S, 0);
        L.v.assign(S, 0);
        L.lv.assign(S, 0);
        if (i == 0) continue;
        for (uint64_t g = 0; g < S; ++g) {
            uint64_t r = rng.next();
            L.ty[g] = (r & 1) ? Mul : Add;
            L.l[g] = (int32_t)((r >> 1) % (uint64_t)i);
            uint64_t r2 = rng.next();
            L.u[g] = (uint32_t)(r2 & (S - 1));
            L.v[g] = (uint32_t)((r2 >> 32) & (S - 1));
        }
    }
    c.inputs.resize(S);
    for (uint64_t g = 0; g < S; ++g) c.inputs[g] = rng.next() & 0x7fffffffULL;  // like random(): < 2^31
    c.subset_init();
    return c;
}

// ------------------------------------------------------------------ challenge stream
GlibcRandom::GlibcRandom(unsigned seed) {
    static_assert(sizeof(buf) >= sizeof(struct random_data), "random_data storage too small");
    memset(buf, 0, sizeof buf);
    memset(state, 0, sizeof state);
    initstate_r(seed, state, sizeof state, (struct random_data*)buf);
}
long GlibcRandom::next() {
    int32_t r;
    random_r((struct random_data*)buf, &r);
    return r;
}
uint64_t GlibcRandom::number() {  // fieldElement.cpp:362-367
    uint64_t ret = (uint64_t)(next() % 10);
    for (int i = 1; i < 20; ++i) ret = (ret * 10ULL + (uint64_t)(next() % 10)) % P;
    return ret;
}
F GlibcRandom::field() {  // fieldElement.cpp:119-124: real first, then img
    F r;
    r.re = number() % P;
    r.im = number() % P;
    return r;
}

ChallengeStream draw_challenges(const Circuit& c, unsigned seed) {
    GlibcRandom rng(seed);
    ChallengeStream cs;
    const int n = c.n_layers();
    const int max_bl = c.max_bit_length();
    auto draw = [&](std::vector<F>& v, int k) {
        v.resize(k);
        for (int i = 0; i < k; ++i) v[i] = rng.field();
        cs.count += (uint64_t)k;
    };
    draw(cs.r_out, c.bit_length(n - 1));                 // verifier.cpp:144-145
    cs.layer.resize(n);
    for (int i = n - 1; i >= 1; --i) {
        LayerChallenges& lc = cs.layer[i];
        draw(lc.r_u, max_bl);                              // verifier.cpp:196 (all max_bl entries)
        lc.assert_random = rng.field();                    // verifier.cpp:202
        ++cs.count;
        int mdb = c.max_dad_bit_length(i);
        if (mdb != -1) draw(lc.r_v, mdb);                  // verifier.cpp:236
        draw(lc.sig, n);                                   // verifier.cpp:278
        draw(lc.r_liu, max_bl);                            // verifier.cpp:279
    }
    return cs;
}

}  // namespace vp
