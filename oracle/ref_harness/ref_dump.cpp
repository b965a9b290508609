// TEST INFRASTRUCTURE ONLY (oracle/_ref build). Runs the UNMODIFIED reference end to end --
// parse + DAG_to_layered (src/main.cpp), subsetInit (src/circuit.cpp), prover (src/prover.cpp),
// verifier incl. the polynomial commitment (src/verifier.cpp, lib/virgo) -- with the logging proxy
// of proxy_prover.h between verifier and prover, and writes:
//   <prefix>.transcript.txt : "TAG real img" per line, emission order
//   <prefix>.circuit.bin    : the reference's layeredCircuit after subsetInit, flat (see dump_circuit)
//   stdout                  : the reference's own statistics lines + "VERIFY 0|1"
// usage: ref_dump <circuit.pws> <out_prefix>
//        ref_dump <circuit.mem> <out_prefix>   a layered circuit handed over directly (gate types, constants and assert flags
//                                              the .pws parser never produces): i32 n_layers; per layer u64 size; per gate
//                                              u8 ty, i32 l, u64 u, u64 v, u64 c.real, u64 c.img, u8 is_assert (layer 0: u = input)
//        REF_TAMPER=<k> in the environment: see proxy_prover.h
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "verifier.h"  // resolved inside the symlink farm -> proxy prover.h

extern layeredCircuit c;
void parse(std::ifstream &circuit_in);
void DAG_to_layered();

static std::vector<std::pair<std::string, F>> g_log;
void ref_log(const char *tag, const F &x) { g_log.emplace_back(tag, x); }
static long g_msg = 0, g_tamper = -1;   // REF_TAMPER: see proxy_prover.h
F ref_tamper(const F &x) { return g_msg++ == g_tamper ? x + F_ONE : x; }

template <class T>
static void put(FILE *f, const T &x) { fwrite(&x, sizeof(T), 1, f); }

// layout: i32 n_layers; per layer: u64 size, i32 bitLength, i32 maxDadBitLength;
//         per gate: u8 ty, i32 l, u64 u, u64 v, u64 lv; per l<i: u64 dadSize, i32 dadBitLength, u64 ids[dadSize]
static void dump_circuit(const char *path) {
    FILE *f = fopen(path, "wb");
    put<int32_t>(f, c.size);
    for (int i = 0; i < c.size; ++i) {
        auto &L = c.circuit[i];
        put<uint64_t>(f, L.size);
        put<int32_t>(f, L.bitLength);
        put<int32_t>(f, L.maxDadBitLength);
        for (u64 g = 0; g < L.size; ++g) {
            auto &G = L.gates[g];
            put<uint8_t>(f, (uint8_t)G.ty);
            put<int32_t>(f, G.l);
            put<uint64_t>(f, G.u);
            put<uint64_t>(f, G.v);
            put<uint64_t>(f, G.lv);
        }
        for (int l = 0; l < i; ++l) {
            put<uint64_t>(f, L.dadSize[l]);
            put<int32_t>(f, L.dadBitLength[l]);
            for (u64 x = 0; x < L.dadSize[l]; ++x) put<uint64_t>(f, L.dadId[l][x]);
        }
    }
    fclose(f);
}

int main(int argc, char **argv) {
    if (argc < 3) {
        fprintf(stderr, "usage: %s <circuit.pws> <out_prefix>\n", argv[0]);
        return 2;
    }
    if (getenv("REF_TAMPER")) g_tamper = atol(getenv("REF_TAMPER"));
    const std::string path = argv[1];
    if (path.size() > 4 && path.substr(path.size() - 4) == ".mem") {
        FILE *m = fopen(argv[1], "rb");
        if (!m) { fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
        auto get = [&](void *dst, size_t n) { if (fread(dst, 1, n, m) != n) { fprintf(stderr, "short circuit file\n"); exit(2); } };
        int32_t n;
        get(&n, 4);
        c.size = n;
        c.circuit.resize(n);
        for (int i = 0; i < n; ++i) {
            layer &L = c.circuit[i];
            uint64_t sz;
            get(&sz, 8);
            L.size = sz;
            L.gates.resize(sz);
            for (u64 g = 0; g < sz; ++g) {
                uint8_t ty, as;
                int32_t l;
                uint64_t u, v, cr, ci;
                get(&ty, 1); get(&l, 4); get(&u, 8); get(&v, 8); get(&cr, 8); get(&ci, 8); get(&as, 1);
                F cc = F_ZERO;
                cc.real = cr;
                cc.img = ci;
                L.gates[g] = i == 0 ? gate(gateType::Input, -1, u, 0, F_ZERO, false) : gate((gateType)ty, l, u, v, cc, as != 0);
            }
            L.bitLength = (int)log2(L.size);   // main.cpp:133-136
            if ((1ULL << L.bitLength) < L.size) ++L.bitLength;
        }
        fclose(m);
    } else {
        std::ifstream in(argv[1]);
        if (!in) { fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
        parse(in);
        DAG_to_layered();
    }
    F::init();
    c.subsetInit();
    dump_circuit((std::string(argv[2]) + ".circuit.bin").c_str());
    prover p(c);
    p.cur_layer = c.size;  // decremented by the proxy on every sumcheckInit
    verifier v(&p, c);
    bool ok = v.verify();
    FILE *f = fopen((std::string(argv[2]) + ".transcript.txt").c_str(), "w");
    for (auto &e : g_log) fprintf(f, "%s %llu %llu\n", e.first.c_str(), e.second.real, e.second.img);
    fclose(f);
    printf("mult counter %d, add counter %d\n", F::multCounter, F::addCounter);
    printf("VERIFY %d\n", ok ? 1 : 0);
    return 0;
}
