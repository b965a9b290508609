#include "proof_io.h"

#include <cstdio>
#include <cstring>

namespace vp {

size_t transcript_len(const Circuit& C) {
    size_t t = 1;
    for (int i = C.n_layers() - 1; i >= 1; --i) {
        const int pb = C.bit_length(i - 1), m = C.max_dad_bit_length(i);
        t += 3 * (size_t)pb + 1;
        if (m != -1) t += 3 * (size_t)m + (size_t)i;
        t += 3 * (size_t)pb + 1;
    }
    return t + 1;
}

namespace {
struct Split {   // where each piece of layer i sits in the flat transcript
    size_t p1, claim_u, p2, claims_v, liu, claim_liu;
    int pb, m;
};
std::vector<Split> split(const Circuit& C, size_t& vres, size_t& input_mle) {
    std::vector<Split> s(C.n_layers());
    size_t t = 0;
    vres = t++;
    for (int i = C.n_layers() - 1; i >= 1; --i) {
        Split& x = s[i];
        x.pb = C.bit_length(i - 1);
        x.m = C.max_dad_bit_length(i);
        x.p1 = t; t += 3 * (size_t)x.pb;
        x.claim_u = t++;
        x.p2 = t;
        x.claims_v = t;
        if (x.m != -1) { t += 3 * (size_t)x.m; x.claims_v = t; t += (size_t)i; }
        x.liu = t; t += 3 * (size_t)x.pb;
        x.claim_liu = t++;
    }
    input_mle = t++;
    return s;
}
void put_u64(std::vector<unsigned char>& o, uint64_t v) {
    unsigned char b[8];
    memcpy(b, &v, 8);
    o.insert(o.end(), b, b + 8);
}
void put_f(std::vector<unsigned char>& o, const F* p, size_t n) {
    const unsigned char* b = reinterpret_cast<const unsigned char*>(p);
    o.insert(o.end(), b, b + n * sizeof(F));
}
}  // namespace

std::vector<unsigned char> transcript_to_gkrproof(const Circuit& C, const F* tr) {
    size_t vres, inp;
    const auto s = split(C, vres, inp);
    const int n = C.n_layers();
    std::vector<unsigned char> o;
    // final_claims_u [layer], final_claims (Liu) [layer]
    put_u64(o, (uint64_t)n);
    for (int i = 0; i < n; ++i) { F z{0, 0}; put_f(o, i ? tr + s[i].claim_u : &z, 1); }
    put_u64(o, (uint64_t)n);
    for (int i = 0; i < n; ++i) { F z{0, 0}; put_f(o, i ? tr + s[i].claim_liu : &z, 1); }
    // final_claims_v [layer][src]
    put_u64(o, (uint64_t)n);
    for (int i = 0; i < n; ++i) {
        const uint64_t len = (i && s[i].m != -1) ? (uint64_t)i : 0;
        put_u64(o, len);
        if (len) put_f(o, tr + s[i].claims_v, len);
    }
    // polys_u, polys_v, polys [layer][round] of quadratic_poly {a,b,c}
    auto polys = [&](int which) {
        put_u64(o, (uint64_t)n);
        for (int i = 0; i < n; ++i) {
            uint64_t len = 0;
            size_t at = 0;
            if (i) {
                if (which == 0) { len = (uint64_t)s[i].pb; at = s[i].p1; }
                if (which == 1) { len = s[i].m != -1 ? (uint64_t)s[i].m : 0; at = s[i].p2; }
                if (which == 2) { len = (uint64_t)s[i].pb; at = s[i].liu; }
            }
            put_u64(o, len);
            if (len) put_f(o, tr + at, 3 * len);
        }
    };
    polys(0);
    polys(1);
    polys(2);
    put_u64(o, 2);
    put_f(o, tr + vres, 1);
    put_f(o, tr + inp, 1);
    return o;
}

std::string gkrproof_to_transcript(const Circuit& C, const unsigned char* b, size_t len, F* tr) {
    size_t vres, inp;
    const auto s = split(C, vres, inp);
    const int n = C.n_layers();
    size_t pos = 0;
    auto get_u64 = [&](uint64_t& v) { if (pos + 8 > len) return false; memcpy(&v, b + pos, 8); pos += 8; return true; };
    auto get_f = [&](F* dst, size_t cnt) { if (pos + cnt * sizeof(F) > len) return false; if (dst) memcpy(dst, b + pos, cnt * sizeof(F)); pos += cnt * sizeof(F); return true; };
    uint64_t cnt;
    for (int which = 0; which < 2; ++which) {   // final_claims_u, final_claims
        if (!get_u64(cnt) || cnt != (uint64_t)n) return "GKRProof: bad claim vector length";
        for (int i = 0; i < n; ++i)
            if (!get_f(i ? tr + (which == 0 ? s[i].claim_u : s[i].claim_liu) : nullptr, 1)) return "GKRProof: truncated";
    }
    if (!get_u64(cnt) || cnt != (uint64_t)n) return "GKRProof: bad final_claims_v length";
    for (int i = 0; i < n; ++i) {
        uint64_t l;
        if (!get_u64(l)) return "GKRProof: truncated";
        const uint64_t want = (i && s[i].m != -1) ? (uint64_t)i : 0;
        if (l != want) return "GKRProof: final_claims_v does not match the circuit";
        if (l && !get_f(tr + s[i].claims_v, l)) return "GKRProof: truncated";
    }
    for (int which = 0; which < 3; ++which) {
        if (!get_u64(cnt) || cnt != (uint64_t)n) return "GKRProof: bad poly vector length";
        for (int i = 0; i < n; ++i) {
            uint64_t l;
            if (!get_u64(l)) return "GKRProof: truncated";
            uint64_t want = 0;
            size_t at = 0;
            if (i) {
                if (which == 0) { want = (uint64_t)s[i].pb; at = s[i].p1; }
                if (which == 1) { want = s[i].m != -1 ? (uint64_t)s[i].m : 0; at = s[i].p2; }
                if (which == 2) { want = (uint64_t)s[i].pb; at = s[i].liu; }
            }
            if (l != want) return "GKRProof: round count does not match the circuit";
            if (l && !get_f(tr + at, 3 * l)) return "GKRProof: truncated";
        }
    }
    if (!get_u64(cnt) || cnt != 2 || !get_f(tr + vres, 1) || !get_f(tr + inp, 1)) return "GKRProof: bad trailer";
    if (pos != len) return "GKRProof: trailing bytes";
    return "";
}

std::string transcript_text(const Circuit& C, const F* tr, const F* ch) {
    size_t vres, inp;
    const auto s = split(C, vres, inp);
    const int n = C.n_layers(), max_bl = C.max_bit_length();
    std::string out;
    char buf[96];
    auto line = [&](const char* tag, const F& x) {
        snprintf(buf, sizeof buf, "%s %llu %llu\n", tag, x.re, x.im);
        out += buf;
    };
    const F zero{0, 0};
    size_t ci = (size_t)C.bit_length(n - 1);
    line("VRES", tr[vres]);
    auto rounds = [&](size_t at, int count, const F* r) {
        F prev = zero;
        for (int j = 0; j < count; ++j) {
            line("CH", prev);
            line("PA", tr[at + 3 * j]);
            line("PB", tr[at + 3 * j + 1]);
            line("PC", tr[at + 3 * j + 2]);
            prev = r[j];
        }
        return prev;
    };
    for (int i = n - 1; i >= 1; --i) {
        const F* r_u = ch + ci;
        ci += (size_t)max_bl + 1;   // r_u + assert_random
        F prev = rounds(s[i].p1, s[i].pb, r_u);
        line("CH", prev);
        line("CLAIM_U", tr[s[i].claim_u]);
        if (s[i].m != -1) {
            const F* r_v = ch + ci;
            ci += (size_t)s[i].m;
            rounds(s[i].p2, s[i].m, r_v);
            for (int l = 0; l < i; ++l) line("CLAIM_V", tr[s[i].claims_v + l]);
        }
        ci += (size_t)n;            // sigma
        const F* r_liu = ch + ci;
        ci += (size_t)max_bl;
        prev = rounds(s[i].liu, s[i].pb, r_liu);
        line("CH", prev);
        line("CLAIM_LIU", tr[s[i].claim_liu]);
    }
    line("INPUT_MLE", tr[inp]);
    return out;
}

}  // namespace vp
