// Test infrastructure: drives the UNMODIFIED reference class `transcriptCache`
// (/root/reference/lib/virgo/src/transcriptCache.hpp:14-50, included from where it lies) with a script read from stdin,
// so that the Fiat-Shamir challenge source of the product (virgo-plus_b200/host/fiat_shamir.h, FsCache) and of the C oracle
// can be compared with the reference's own class on the same sequence of stores and draws.
//   script, one operation per line:   S <hex bytes>   -> store(ptr, n)        (transcriptCache.hpp:18-23)
//                                     F <re> <im>     -> store(fieldElement)  (:25-31; 16 bytes {real, img})
//                                     R               -> random()             (:40-46); prints "<real> <img>"
#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "virgo/src/fieldElement.hpp"   // brings typedef.hpp's u64 / i64, which transcriptCache.hpp uses
#include "virgo/src/transcriptCache.hpp"

int main() {
    virgo::fieldElement::init();
    static transcriptCache tc;   // 100 kB pool: not on the stack
    std::string line;
    while (std::getline(std::cin, line)) {
        if (line.empty()) continue;
        std::istringstream is(line);
        std::string op;
        is >> op;
        if (op == "S") {
            std::string hex;
            is >> hex;
            std::vector<unsigned char> b(hex.size() / 2);
            for (size_t i = 0; i < b.size(); ++i) b[i] = (unsigned char)strtoul(hex.substr(2 * i, 2).c_str(), nullptr, 16);
            tc.store(b.data(), b.size());
        } else if (op == "F") {
            unsigned long long re, im;
            is >> re >> im;
            virgo::fieldElement x;
            x.real = re;
            x.img = im;
            tc.store(x);
        } else if (op == "R") {
            virgo::fieldElement r = tc.random();
            printf("%llu %llu\n", (unsigned long long)r.real, (unsigned long long)r.img);
        }
    }
    return 0;
}
