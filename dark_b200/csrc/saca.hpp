// saca.hpp — C++ host-side mirror of the reference's `saca` module over the C ABI.
//
// The reference host code is Rust (no Rust toolchain in this image), so the compiled-language
// mirror of `saca::Constructor` (/root/reference/src/saca.rs:344-384) is this header; the Rust
// wrapper in rust/ is the same thing in the reference's own language (uncompiled here).
//   Symbol, Suffix, SUF_INVALID                        saca.rs:18-22
//   Constructor(max_n) / capacity / compute / reuse    saca.rs:351-383
//   bwt(input) -> (bytes, origin): replaces compute + compress::bwt::TransformIterator at
//                                                      block/dc.rs:45-50 and block/raw.rs:39-44
// Failures throw std::runtime_error where the Rust code panics.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/dark_bwt.h"

namespace dark {
namespace saca {

using Symbol = uint8_t;
using Suffix = uint32_t;
constexpr Suffix SUF_INVALID = ~Suffix(0);

class Constructor {
public:
    explicit Constructor(size_t max_n, int device = 0) : n_(max_n) {
        check(dark_bwt_create(max_n, device, &ctx_), "Constructor::new");
    }
    ~Constructor() { dark_bwt_destroy(ctx_); }
    Constructor(const Constructor&) = delete;
    Constructor& operator=(const Constructor&) = delete;

    size_t capacity() const { return (size_t)dark_bwt_capacity(ctx_); }

    // compute: input.len() must equal capacity (assert_eq!, saca.rs:369).  Returns the SA, valid
    // until the next call (the `&'a [Suffix]` borrow of the reference).
    const Suffix* compute(const Symbol* input, size_t len) {
        if (len != n_) throw std::runtime_error("assertion failed: input.len() == self.n");
        sa_.resize(len);
        bwt_.resize(len);
        uint64_t origin = 0;
        check(dark_bwt_forward(ctx_, input, len, bwt_.data(), &origin, sa_.data(), &stats_), "Constructor::compute");
        return sa_.data();
    }

    // the fused call-site entry: BWT bytes + origin, no SA on the host
    std::pair<std::vector<Symbol>, size_t> bwt(const Symbol* input, size_t len) {
        std::vector<Symbol> out(len);
        uint64_t origin = 0;
        check(dark_bwt_forward(ctx_, input, len, out.data(), &origin, nullptr, &stats_), "Constructor::bwt");
        return {std::move(out), (size_t)origin};
    }

    // reuse: the context's host scratch (>= capacity words)
    std::pair<Suffix*, size_t> reuse() {
        uint32_t* p = nullptr;
        uint64_t cnt = 0;
        check(dark_bwt_reuse(ctx_, &p, &cnt), "Constructor::reuse");
        return {p, (size_t)cnt};
    }

    // Many small blocks in one pass over the GPU (dark_bwt_forward_many): (bwt, origin) per block, each identical to
    // bwt(block).  The blocks share the arena: their lengths must sum to <= capacity().
    std::vector<std::pair<std::vector<Symbol>, size_t>> bwt_many(const std::vector<std::pair<const Symbol*, size_t>>& blocks) {
        const size_t cnt = blocks.size();
        std::vector<std::pair<std::vector<Symbol>, size_t>> res(cnt);
        std::vector<const uint8_t*> texts(cnt);
        std::vector<uint64_t> ns(cnt), origins(cnt);
        std::vector<uint8_t*> outs(cnt);
        for (size_t k = 0; k < cnt; ++k) {
            texts[k] = blocks[k].first;
            ns[k] = blocks[k].second;
            res[k].first.resize(blocks[k].second);
            outs[k] = res[k].first.data();
        }
        check(dark_bwt_forward_many(ctx_, texts.data(), ns.data(), outs.data(), origins.data(), cnt, &stats_), "Constructor::bwt_many");
        for (size_t k = 0; k < cnt; ++k) res[k].second = (size_t)origins[k];
        return res;
    }

    const dark_bwt_stats& stats() const { return stats_; }
    dark_bwt_ctx* raw() { return ctx_; }

private:
    void check(int rc, const char* what) const {
        if (rc != DARK_BWT_OK) {
            std::string msg = std::string(what) + ": " + dark_bwt_strerror(rc);
            if (ctx_ && dark_bwt_last_error(ctx_)[0]) msg += std::string(" [") + dark_bwt_last_error(ctx_) + "]";
            throw std::runtime_error(msg);
        }
    }
    dark_bwt_ctx* ctx_ = nullptr;
    size_t n_;
    std::vector<Suffix> sa_;
    std::vector<Symbol> bwt_;
    dark_bwt_stats stats_{};
};

}  // namespace saca
}  // namespace dark
