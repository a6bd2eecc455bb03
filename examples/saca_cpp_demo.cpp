// C++ host-side use of the drop-in boundary through the mirror of saca::Constructor
// (dark_b200/csrc/saca.hpp).  Mirrors the reference's own test `saca::test::some_detail`
// (/root/reference/src/saca.rs:393-407) on its two known-answer vectors (saca.rs:411-412).
//
//   g++ -std=c++17 -I. examples/saca_cpp_demo.cpp -Ldark_b200/lib -ldark_bwt -Wl,-rpath,$PWD/dark_b200/lib -o saca_cpp_demo
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "dark_b200/csrc/saca.hpp"

static int check(const char* text, const std::vector<uint32_t>& sa_expected, size_t origin_expected, const char* bwt_expected) {
    const size_t n = strlen(text);
    dark::saca::Constructor con(n);
    if (con.capacity() != n) return 1;
    const dark::saca::Suffix* suf = con.compute(reinterpret_cast<const uint8_t*>(text), n);
    for (size_t i = 0; i < n; ++i)
        if (suf[i] != sa_expected[i]) return 2;
    auto out = con.bwt(reinterpret_cast<const uint8_t*>(text), n);
    if (out.second != origin_expected) return 3;
    if (memcmp(out.first.data(), bwt_expected, n) != 0) return 4;
    auto scratch = con.reuse();
    if (scratch.second < n) return 5;
    std::printf("%s: SA, origin %zu and BWT \"%s\" match; %u rounds, %.3f ms on the device\n", text, out.second, bwt_expected,
                con.stats().rounds, con.stats().device_ms);
    return 0;
}

int main() {
    try {
        int rc = check("abracadabra", {10, 7, 0, 3, 5, 8, 1, 4, 6, 9, 2}, 2, "rdarcaaaabb");
        if (rc) return rc;
        rc = check("banana", {5, 3, 1, 0, 4, 2}, 3, "nnbaaa");
        if (rc) return 10 + rc;
        {  // both known answers again, as two blocks sorted in one pass (dark_bwt_forward_many)
            dark::saca::Constructor many(64);
            auto res = many.bwt_many({{reinterpret_cast<const uint8_t*>("abracadabra"), 11}, {reinterpret_cast<const uint8_t*>("banana"), 6}});
            if (res.size() != 2 || res[0].second != 2 || res[1].second != 3) return 30;
            if (memcmp(res[0].first.data(), "rdarcaaaabb", 11) != 0 || memcmp(res[1].first.data(), "nnbaaa", 6) != 0) return 31;
        }
        try {  // n == 1 panics in the reference (saca.rs:300): the mirror throws
            dark::saca::Constructor one(2);
            one.bwt(reinterpret_cast<const uint8_t*>("x"), 1);
            return 20;
        } catch (const std::runtime_error&) {
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 99;
    }
    std::puts("ok");
    return 0;
}
